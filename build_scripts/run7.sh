set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_part1.py tests/test_gpu_golden.py -m gpu -x -q -k "linkage or seams" > gpurun_out/r2_t7.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; echo "rc=$?" >> gpurun_out/r2_b3.err
HICA_NN_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/nn_dbg4.json 2> gpurun_out/nn_dbg4.err
HICA_NN_KERNEL=3 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b3_nn3.json 2> /dev/null
HICA_NN_FLAGS=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b3_nopf.json 2> /dev/null
