set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_part1.py tests/test_gpu_golden.py -m gpu -x -q -k "linkage or seams" > gpurun_out/r2_t12.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t12.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b12_ilv.json 2> gpurun_out/r2_b12_ilv.err; echo "rc=$?" >> gpurun_out/r2_b12_ilv.err
HICA_NN_FLAGS=4 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b12_contig.json 2> gpurun_out/r2_b12_contig.err
HICA_NN_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/nn_dbg12.json 2> gpurun_out/nn_dbg12.err
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --part1-only --bins 32000 --no-golden-check > gpurun_out/r2_b12_50kb.json 2> gpurun_out/r2_b12_50kb.err
