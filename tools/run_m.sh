set -x
mkdir -p gpurun_out
T=${TAG:-r02w}
echo "== default build (512 threads x 2 CTAs, stage 960, 8 loads in flight)" > gpurun_out/${T}_ab.log
timeout 900 python -m pytest tests/test_gpu_evrep_tiled.py -x -q 2>&1 | tail -2 >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_evrep.py --only-tiled >> gpurun_out/${T}_ab.log 2>&1
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
for V in "-DEP_EV_THREADS=384 -DEP_EV_CTAS=3 -DEP_EV_STAGE=512" "-DEP_EV_THREADS=256 -DEP_EV_CTAS=4 -DEP_EV_STAGE=512" "-DEP_EV_THREADS=1024 -DEP_EV_CTAS=1 -DEP_EV_STAGE=960"; do
  echo "== $V" >> gpurun_out/${T}_ab.log
  EP_NVCC_EXTRA="$V" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
  timeout 300 python tools/quick_evrep.py --check >> gpurun_out/${T}_ab.log 2>&1
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_ab.log
