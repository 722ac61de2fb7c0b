set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t23.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t23.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b23.json 2> gpurun_out/r2_b23.err; echo "rc=$?" >> gpurun_out/r2_b23.err
HICA_ROWSUM_DEFER=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b23_nodefer.json 2> gpurun_out/r2_b23_nodefer.err
