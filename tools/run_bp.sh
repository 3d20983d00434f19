set -x
mkdir -p gpurun_out
T=r02zf
timeout 600 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_tiled.py -q -m gpu 2>&1 | tail -3 > gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
cat gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_1gpu.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print(json.dumps(d['e2e']['from_reference_format'],indent=0))
print(d['e2e']['value'])
"
