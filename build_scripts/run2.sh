set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ties.py tests/test_gpu_part1.py tests/test_gpu_golden.py tests/test_gpu_files.py -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; echo "rc=$?" >> gpurun_out/r2_b2.err
