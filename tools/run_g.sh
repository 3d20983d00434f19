set -x
mkdir -p gpurun_out
T=${TAG:-r02o}
timeout 900 python -m pytest tests/test_gpu_evrep_tiled.py -x -q 2>&1 | tail -30 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
