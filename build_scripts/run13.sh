set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_part1.py tests/test_gpu_golden.py -m gpu -x -q -k "linkage or seams" > gpurun_out/r2_t13.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t13.log
for f in 0 4 8 12; do
HICA_NN_FLAGS=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b13_f$f.json 2> gpurun_out/r2_b13_f$f.err; echo "rc=$?" >> gpurun_out/r2_b13_f$f.err
done
HICA_NN_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/nn_dbg13_f0.json 2> gpurun_out/nn_dbg13_f0.err
HICA_NN_FLAGS=4 HICA_NN_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/nn_dbg13_f4.json 2> gpurun_out/nn_dbg13_f4.err
