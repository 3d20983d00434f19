set -x
mkdir -p gpurun_out
T=r02zh
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/${T}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
cat gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log
