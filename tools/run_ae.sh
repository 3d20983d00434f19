set -x
mkdir -p gpurun_out
T=${TAG:-r02ar}
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
: > gpurun_out/${T}_ab.log
for V in "-DEP_ABL_EVREP_C" "-DEP_ABL_EVREP_C -DEP_ABL_EVREP_STORE"; do
  echo "== EvRep ablation $V (wrong results by construction)" >> gpurun_out/${T}_ab.log
  EP_NVCC_EXTRA="$V" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
  timeout 300 python tools/quick_evrep.py --only-tiled >> gpurun_out/${T}_ab.log 2>&1
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_ab.log
