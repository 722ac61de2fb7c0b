set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden_large.py tests/test_launcher_reference.py tests/test_gpu_launcher.py -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_golden_large.py --deselect tests/test_launcher_reference.py > gpurun_out/r2_t2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo "rc=$?" >> gpurun_out/r2_b1.err
timeout 300 python bench.py --config candidates --steps 5 --warmup 3 > gpurun_out/r2_c1.json 2> gpurun_out/r2_c1.err; echo "rc=$?" >> gpurun_out/r2_c1.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_ref.json 2> gpurun_out/r2_ref.err; echo "rc=$?" >> gpurun_out/r2_ref.err
