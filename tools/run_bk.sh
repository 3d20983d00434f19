set -x
mkdir -p gpurun_out
T=r02bk
timeout 900 python -m pytest tests/test_gpu_tiled.py -q -x 2>&1 | tail -5 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 10 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.log
timeout 300 python tools/quick_c5.py 2>&1 | tail -7 | tee gpurun_out/${T}_c5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane$' -c 1 -f -o gpurun_out/${T}_plane_c5 python tools/quick_c5.py > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
