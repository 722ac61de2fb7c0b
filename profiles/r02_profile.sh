# Round-2 profiling pass (one B200): launch list of one step + ncu --set full captures of the round's kernels.
# Every profiled command is first run WITHOUT ncu; numbers printed under ncu are never bench values.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 850 -c 900 --csv --log-file gpurun_out/r02_launches_100kb.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-golden-check > gpurun_out/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_nn_chain4' -c 1 -f -o gpurun_out/r02_nn4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --part1-only --no-golden-check > gpurun_out/r02_ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rowsum_pairwise|k_build_W|k_rowsum_sequential|k_nn_init' -c 4 -f -o gpurun_out/r02_k1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --part1-only --no-golden-check > gpurun_out/r02_ncu_f2.log 2>&1
python bench.py --config candidates --steps 2 --warmup 1 > gpurun_out/r02_cand_plain.json 2> gpurun_out/r02_cand_plain.err
ncu --set full --clock-control none --import-source on -k regex:'k_score' -c 1 -f -o gpurun_out/r02_k_score python bench.py --config candidates --steps 1 --warmup 1 > gpurun_out/r02_ncu_f3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
