set -x
mkdir -p gpurun_out
T=r02bi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_view_bilinear_staged' -c 1 -f -o gpurun_out/${T}_view python tools/quick_c5.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
