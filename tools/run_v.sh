set -x
mkdir -p gpurun_out
T=${TAG:-r02ag}
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
echo "== -DEP_SWEEP_NORETURN (experiment: fire-and-forget shared atomics, no wrap detection)" > gpurun_out/${T}_ab.log
EP_NVCC_EXTRA="-DEP_SWEEP_NORETURN" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
timeout 300 python tools/quick_bin.py --batch 256 --packed4 --methods tiled,global --check 2>&1 | grep -v "^global" >> gpurun_out/${T}_ab.log
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_ab.log
