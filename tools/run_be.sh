set -x
mkdir -p gpurun_out
T=r02be
timeout 900 python -m pytest tests/test_gpu_tiled.py -q -x 2>&1 | tail -15 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
echo "== default (1024 threads, register pipeline, y by multiply-high)" | tee -a gpurun_out/${T}_ab.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane,tiled --check --steps 10 2>&1 | tail -4 | tee -a gpurun_out/${T}_ab.log
echo "== EP_PLANE_LUT=1" | tee -a gpurun_out/${T}_ab.log
EP_PLANE_LUT=1 timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 10 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.log
echo "== 240x180 unscaled-like (scale 240/640) and 256x256 (scale 0.4: both axes multiply-high?)" | tee -a gpurun_out/${T}_ab.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 192x256 --methods plane,tiled --check --steps 10 2>&1 | tail -4 | tee -a gpurun_out/${T}_ab.log
for V in t512; do
  cp build/variants/$V.so eventpretrain_b200/libeventpretrain_b200.so
  echo "== $V" | tee -a gpurun_out/${T}_ab.log
  timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 10 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.log
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane$' -c 1 -f -o gpurun_out/${T}_plane python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
