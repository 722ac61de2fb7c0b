set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard32.py -m gpu -x -q > gpurun_out/r2_t6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t6.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 10kb --bins 64000 --steps 2 --warmup 1 > gpurun_out/r2_10kb_2gpu_64k.json 2> gpurun_out/r2_10kb_2gpu_64k.err; echo "rc=$?" >> gpurun_out/r2_10kb_2gpu_64k.err
