set -x
mkdir -p gpurun_out
T=${TAG:-r02v}
echo "== default build (RPL=4, depth 2)" > gpurun_out/${T}_ab.log
timeout 900 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -2 >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --size 224x224 >> gpurun_out/${T}_ab.log 2>&1
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
for V in "-DEP_ITEM_RPL=6 -DEP_ITEM_DEPTH=1" "-DEP_ITEM_RPL=5 -DEP_ITEM_DEPTH=2" "-DEP_ITEM_RPL=6 -DEP_ITEM_DEPTH=2" "-DEP_ITEM_RPL=8 -DEP_ITEM_DEPTH=1"; do
  echo "== $V" >> gpurun_out/${T}_ab.log
  EP_NVCC_EXTRA="$V" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
  timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
  timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --size 224x224 >> gpurun_out/${T}_ab.log 2>&1
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_ab.log
