set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard32.py tests/test_gpu_ties.py -m gpu -x -q > gpurun_out/r2_t4.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t4.log
timeout 600 python bench.py --config 10kb --bins 24000 --steps 2 --warmup 1 > gpurun_out/r2_10kb_small.json 2> gpurun_out/r2_10kb_small.err; echo "rc=$?" >> gpurun_out/r2_10kb_small.err
