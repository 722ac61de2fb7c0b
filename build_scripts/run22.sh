set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b22.json 2> gpurun_out/r2_b22.err; echo "rc=$?" >> gpurun_out/r2_b22.err
HICA_P2_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r2_p2prof22.json 2> gpurun_out/r2_p2prof22.err
