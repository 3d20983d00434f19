set -x
mkdir -p gpurun_out
T=${TAG:-r02x}
echo "== default build (512 x 2, 12800 cells)" > gpurun_out/${T}_ab.log
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled >> gpurun_out/${T}_ab.log 2>&1
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
for V in "-DEP_SWEEP_THREADS=320 -DEP_SWEEP_CTAS=3 -DEP_TILE_CELLS=7680" "-DEP_SWEEP_THREADS=256 -DEP_SWEEP_CTAS=4 -DEP_TILE_CELLS=5760" "-DEP_SWEEP_THREADS=384 -DEP_SWEEP_CTAS=3 -DEP_TILE_CELLS=7680" "-DEP_SWEEP_THREADS=1024 -DEP_SWEEP_CTAS=1 -DEP_TILE_CELLS=16000"; do
  echo "== $V" >> gpurun_out/${T}_ab.log
  EP_NVCC_EXTRA="$V" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
  timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
  timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled --skewed >> gpurun_out/${T}_ab.log 2>&1
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_ab.log
