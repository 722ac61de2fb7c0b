set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --config 10kb --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02b_10kb_2gpu.json 2> gpurun_out/r02b_10kb_2gpu.err; echo "rc=$?" >> gpurun_out/r02b_10kb_2gpu.err
