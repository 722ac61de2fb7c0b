set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_2gpu_smi.txt
timeout 900 python -m pytest tests/test_gpu_shard32.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_t5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t5.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 10kb --bins 48000 --steps 2 --warmup 1 > gpurun_out/r2_10kb_2gpu_small.json 2> gpurun_out/r2_10kb_2gpu_small.err; echo "rc=$?" >> gpurun_out/r2_10kb_2gpu_small.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_2gpu.json 2> gpurun_out/r2_b_2gpu.err; echo "rc=$?" >> gpurun_out/r2_b_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config candidates --steps 5 --warmup 3 > gpurun_out/r2_c_2gpu.json 2> gpurun_out/r2_c_2gpu.err; echo "rc=$?" >> gpurun_out/r2_c_2gpu.err
