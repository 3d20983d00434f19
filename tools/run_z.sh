set -x
mkdir -p gpurun_out
T=${TAG:-r02ak}
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
: > gpurun_out/${T}_ab.log
for V in "-DEP_ABL_ROUTE_LUT" "-DEP_ABL_ROUTE_ATOM" "-DEP_ABL_ROUTE_STAGE" "-DEP_ABL_ROUTE_OUT" "-DEP_ABL_ROUTE_LUT -DEP_ABL_ROUTE_ATOM -DEP_ABL_ROUTE_STAGE" "-DEP_ABL_ROUTE_LUT -DEP_ABL_ROUTE_ATOM -DEP_ABL_ROUTE_STAGE -DEP_ABL_ROUTE_OUT"; do
  echo "== route ablation $V (results are wrong by construction; look at pass1)" >> gpurun_out/${T}_ab.log
  EP_NVCC_EXTRA="$V" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
  timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled 2>&1 | tail -1 >> gpurun_out/${T}_ab.log
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_ab.log
