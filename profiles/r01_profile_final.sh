set -x
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2400 -c 800 --csv --log-file gpurun_out/r01_launches_100kb_final2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_nn_chain3|k_rank_radix32|k_rank_remap' -c 3 -f -o gpurun_out/r01_final2_p1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ra_counts' -c 1 -f -o gpurun_out/r01_final2_cuts python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_g_tables|k_g_exact_near' --launch-skip 6 -c 2 -f -o gpurun_out/r01_final2_p2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
