set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
HICA_LAUNCHER_RECORD=$PWD/gpurun_out/r02_launcher_500kb.json timeout 900 python -m pytest tests/test_launcher_reference.py -m gpu -x -q -s -k 500kb > gpurun_out/r2_t10.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t10.log
timeout 1200 python profiles/launcher_100kb.py gpurun_out/r02_launcher_100kb.json > gpurun_out/r2_l100.log 2>&1; echo "rc=$?" >> gpurun_out/r2_l100.log
