set -x
mkdir -p gpurun_out
T=${TAG:-r02ao}
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_guards.py -x -q 2>&1 | tail -3 > gpurun_out/${T}_ab.log
echo "== default (auto)" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
for D in 0 1; do
for BINS in 9 15; do
  echo "== EP_DEFERRED_SUM=$D bins $BINS batch 64" >> gpurun_out/${T}_ab.log
  EP_DEFERRED_SUM=$D timeout 300 python tools/quick_bin.py --batch 64 --packed4 --methods tiled,global --check --bins $BINS 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
done
done
echo "== EP_DEFERRED_SUM=0 bins 5" >> gpurun_out/${T}_ab.log
EP_DEFERRED_SUM=0 timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
