set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_part1.py tests/test_gpu_golden.py -m gpu -x -q -k "linkage or seams" > gpurun_out/r2_t15.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t15.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b15_nn4.json 2> gpurun_out/r2_b15_nn4.err; echo "rc=$?" >> gpurun_out/r2_b15_nn4.err
HICA_NN_FLAGS=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b15_nn4f1.json 2> gpurun_out/r2_b15_nn4f1.err
HICA_NN_KERNEL=3 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b15_nn3.json 2> gpurun_out/r2_b15_nn3.err
HICA_NN_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/nn_dbg15.json 2> gpurun_out/nn_dbg15.err
