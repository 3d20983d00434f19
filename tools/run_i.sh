set -x
mkdir -p gpurun_out
T=${TAG:-r02q}
timeout 600 python -m pytest tests/test_gpu_evrep_tiled.py tests/test_gpu_stage1.py -x -q 2>&1 | tail -5 > gpurun_out/${T}_tests.log
timeout 300 python tools/quick_evrep.py --check > gpurun_out/${T}_evrep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_route|k_evrep_sweep" -c 2 -f -o gpurun_out/${T}_evrep python tools/quick_evrep.py --batch 8 --steps 1 --only-tiled > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_tests.log gpurun_out/${T}_evrep.log
