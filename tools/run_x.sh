set -x
mkdir -p gpurun_out
T=${TAG:-r02ai}
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_guards.py -x -q 2>&1 | tail -12 > gpurun_out/${T}_ab.log
echo "== paired halves with flush permit (default)" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --skewed >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --size 224x224 >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --stats >> gpurun_out/${T}_ab.log 2>&1
echo "== EP_SWEEP_PAIR=0" >> gpurun_out/${T}_ab.log
EP_SWEEP_PAIR=0 timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
