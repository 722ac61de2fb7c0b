# Round-end check on one B200: the GPU test suite, smoke(), the default bench line and the reference arm.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?" >> gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "rc=$?" >> gpurun_out/final_bench_reference.err
# launch list of the final build (after the plain run above exited 0; numbers printed under ncu are never bench values)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-golden-check > gpurun_out/final_ncu_l.log 2>&1; echo "rc=$?" >> gpurun_out/final_ncu_l.log
HICA_FB_CHUNK=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b_chunk1.json 2> gpurun_out/r2_b_chunk1.err; echo "rc=$?" >> gpurun_out/r2_b_chunk1.err
