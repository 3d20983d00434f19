set -x
mkdir -p gpurun_out
T=${TAG:-r02p}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
tail -5 gpurun_out/${T}_bench_1gpu.err
cat gpurun_out/${T}_tests.log
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_1gpu.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print(json.dumps(d['roofline'],indent=0)[:1500])
print(json.dumps(d['e2e'],indent=0)[:2500])
print(json.dumps(d['cpu_baseline']))
print(json.dumps(d['extra'],indent=0)[:7000])
"
