set -x
mkdir -p gpurun_out
T=${TAG:-r02al}
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_guards.py -x -q 2>&1 | tail -12 > gpurun_out/${T}_ab.log
echo "== deferred sum (default)" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --skewed >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check --size 224x224 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --stats >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 32 --packed4 --methods tiled,global --check --bins 15 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
cat gpurun_out/${T}_ab.log
