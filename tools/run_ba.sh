set -x
mkdir -p gpurun_out
T=r02ba
timeout 900 python -m pytest tests/test_gpu_tiled.py -q -x 2>&1 | tail -15 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods tiled,plane,global --check --steps 10 2>&1 | tail -8 | tee gpurun_out/${T}_224.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --skewed --steps 10 2>&1 | tail -3 | tee -a gpurun_out/${T}_224.log
