set -x
mkdir -p gpurun_out
T=r02bf
timeout 900 python -m pytest tests/test_gpu_tiled.py -q -x 2>&1 | tail -15 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 10 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
tail -3 gpurun_out/${T}_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_1gpu.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print(json.dumps(d['extra']['reference_res_224'],indent=0))
print(json.dumps(d['extra']['configs'],indent=0)[:6000])
"
