set -x
mkdir -p gpurun_out
T=r02bm
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_views.py -q -x -m gpu 2>&1 | tail -8 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane,tiled --steps 10 2>&1 | tail -2 | tee -a gpurun_out/${T}_ab.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane,tiled --stats --steps 10 2>&1 | tail -2 | tee -a gpurun_out/${T}_ab.log
