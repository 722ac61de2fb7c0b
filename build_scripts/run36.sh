set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard32.py -m gpu -x -q > gpurun_out/r2_t36.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t36.log
