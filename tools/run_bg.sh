set -x
mkdir -p gpurun_out
T=r02bg
timeout 300 python tools/quick_c5.py 2>&1 | tail -6 | tee gpurun_out/${T}_c5.log
