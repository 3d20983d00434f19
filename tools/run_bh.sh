set -x
mkdir -p gpurun_out
T=r02bh
timeout 900 python -m pytest tests/test_views.py -q -x -m gpu 2>&1 | tail -15 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 300 python tools/quick_c5.py 2>&1 | tail -6 | tee gpurun_out/${T}_c5.log
