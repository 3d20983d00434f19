set -x
mkdir -p gpurun_out
T=${TAG:-r02u}
timeout 900 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -3 > gpurun_out/${T}_tests.log
echo "== arith route (default)" > gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --skewed >> gpurun_out/${T}_ab.log 2>&1
echo "== LUT route (EP_ROUTE_ARITH=0)" >> gpurun_out/${T}_ab.log
EP_ROUTE_ARITH=0 timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
echo "== 64-record items (-DEP_ITEM_RPL=2)" >> gpurun_out/${T}_ab.log
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
EP_NVCC_EXTRA="-DEP_ITEM_RPL=2" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --size 224x224 >> gpurun_out/${T}_ab.log 2>&1
echo "== 96-record items (-DEP_ITEM_RPL=3)" >> gpurun_out/${T}_ab.log
EP_NVCC_EXTRA="-DEP_ITEM_RPL=3" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check >> gpurun_out/${T}_ab.log 2>&1
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_tests.log gpurun_out/${T}_ab.log
