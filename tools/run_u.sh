set -x
mkdir -p gpurun_out
T=${TAG:-r02af}
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_guards.py -x -q 2>&1 | tail -15 > gpurun_out/${T}_ab.log
echo "== all-planes sweep (default)" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --skewed >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check --size 224x224 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
echo "== EP_SWEEP_ALL=0" >> gpurun_out/${T}_ab.log
EP_SWEEP_ALL=0 timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
