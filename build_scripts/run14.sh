set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in 256 384 512; do
HICA_NN_THREADS=$t timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b14_t$t.json 2> gpurun_out/r2_b14_t$t.err; echo "rc=$?" >> gpurun_out/r2_b14_t$t.err
done
HICA_NN_CLUSTER=8 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b14_c8.json 2> gpurun_out/r2_b14_c8.err
HICA_RANK_OVERLAP=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --part1-only > gpurun_out/r2_b14_noov.json 2> gpurun_out/r2_b14_noov.err
