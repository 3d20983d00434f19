set -x
mkdir -p gpurun_out
T=${TAG:-r02ae}
timeout 900 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -2 > gpurun_out/${T}_ab.log
echo "== predicated atomics (default build)" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --skewed >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --size 224x224 >> gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
