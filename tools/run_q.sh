set -x
mkdir -p gpurun_out
T=${TAG:-r02aa}
echo "== tests (default G)" > gpurun_out/${T}_ab.log
timeout 900 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -2 >> gpurun_out/${T}_ab.log
echo "== tests with EP_TILED_GROUPS=4" >> gpurun_out/${T}_ab.log
EP_TILED_GROUPS=4 timeout 900 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -2 >> gpurun_out/${T}_ab.log
for CFG in "1 1 2" "4 1 2" "8 1 2" "16 1 2" "8 2 2" "8 1 1" "16 2 2" "32 1 2"; do
  set -- $CFG
  echo "== groups $1 route_ctas $2 sweep_ctas $3" >> gpurun_out/${T}_ab.log
  EP_TILED_GROUPS=$1 EP_TILED_ROUTE_CTAS=$2 EP_TILED_SWEEP_CTAS=$3 timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
done
cat gpurun_out/${T}_ab.log
