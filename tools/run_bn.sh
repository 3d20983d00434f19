set -x
mkdir -p gpurun_out
T=r02zd
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/${T}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
cat gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_1gpu.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print(json.dumps(d['extra']['reference_res_224'],indent=0))
"
