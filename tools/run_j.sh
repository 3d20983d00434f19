set -x
mkdir -p gpurun_out
T=${TAG:-r02s}
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_evrep_tiled.py tests/test_gpu_stage1.py -x -q 2>&1 | tail -5 > gpurun_out/${T}_tests.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled > gpurun_out/${T}_q256.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --size 224x224 > gpurun_out/${T}_q256_224.log 2>&1
timeout 300 python tools/quick_evrep.py --only-tiled > gpurun_out/${T}_evrep.log 2>&1
cat gpurun_out/${T}_tests.log gpurun_out/${T}_q256.log gpurun_out/${T}_q256_224.log gpurun_out/${T}_evrep.log
