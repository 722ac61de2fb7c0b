set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 10kb --steps 2 --warmup 1 > gpurun_out/r2_10kb_${N}gpu.json 2> gpurun_out/r2_10kb_${N}gpu.err; echo "rc=$?" >> gpurun_out/r2_10kb_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config candidates --steps 5 --warmup 3 > gpurun_out/r2_c_${N}gpu.json 2> gpurun_out/r2_c_${N}gpu.err; echo "rc=$?" >> gpurun_out/r2_c_${N}gpu.err
