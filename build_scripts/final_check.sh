# Round-end check on one B200: the GPU test suite, smoke(), the default bench line and the reference arm.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?" >> gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "rc=$?" >> gpurun_out/final_bench_reference.err
# The launch list / ncu captures are a separate pass (profiles/r02_profile.sh): under ncu the synthetic-workload
# generator's ~5,000 torch kernels are profiled too unless --launch-skip jumps over them, which takes minutes.
