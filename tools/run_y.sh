set -x
mkdir -p gpurun_out
T=${TAG:-r02aj}
echo "== pair, no permit" > gpurun_out/${T}_ab.log
EP_SWEEP_PERMIT=0 timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
echo "== pair, permit" >> gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
