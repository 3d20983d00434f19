set -x
mkdir -p gpurun_out
T=${TAG:-r02aq}
timeout 900 python -m pytest tests/test_gpu_evrep_tiled.py tests/test_gpu_guards.py tests/test_gpu_smoke.py -x -q 2>&1 | tail -12 > gpurun_out/${T}_ab.log
timeout 300 python tools/quick_evrep.py --check >> gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
