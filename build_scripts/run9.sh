set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_part2.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py tests/test_gpu_ties.py -m gpu -x -q > gpurun_out/r2_t9.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t9.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err; echo "rc=$?" >> gpurun_out/r2_b5.err
timeout 300 python bench.py --bins 32000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_b5_50kb.json 2> gpurun_out/r2_b5_50kb.err; echo "rc=$?" >> gpurun_out/r2_b5_50kb.err
