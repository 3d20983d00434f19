set -x
mkdir -p gpurun_out
T=r02bo
timeout 900 python -m pytest tests/test_gpu_tiled.py -q -x 2>&1 | tail -6 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane,tiled --check --steps 20 2>&1 | tail -4 | tee -a gpurun_out/${T}_ab.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --skewed --steps 20 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.log
timeout 300 python tools/quick_c5.py 2>&1 | tail -7 | grep "bin plane" | tee -a gpurun_out/${T}_ab.log
