set -x
mkdir -p gpurun_out
T=${TAG:-r02m}
timeout 900 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -5 > gpurun_out/${T}_tests.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check > gpurun_out/${T}_q256.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --stats > gpurun_out/${T}_q256_stats.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --check --skewed > gpurun_out/${T}_q256_skew.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check --size 224x224 > gpurun_out/${T}_q256_224.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_route|k_sweep" -c 2 -f -o gpurun_out/${T}_tiled python tools/quick_bin.py --batch 64 --packed4 --methods tiled --steps 1 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_tests.log gpurun_out/${T}_q256.log gpurun_out/${T}_q256_stats.log gpurun_out/${T}_q256_skew.log gpurun_out/${T}_q256_224.log
