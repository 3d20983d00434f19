set -x
mkdir -p gpurun_out
T=${TAG:-r02l}
timeout 600 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -4 > gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
tail -3 gpurun_out/${T}_bench_1gpu.err
# DRAM traffic of the headline step: few metrics, caches NOT flushed between kernels (what the kernels really see back to back)
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:"k_route|k_sweep|k_stats|k_sample_meta|k_tiled_setup" -c 20 --csv --log-file gpurun_out/${T}_traffic.csv python tools/quick_bin.py --batch 256 --packed4 --methods tiled --steps 1 > gpurun_out/${T}_traffic.log 2>&1
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1
cat gpurun_out/${T}_tests.log
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_1gpu.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches')}))
print(d['roofline']['frac'], d['roofline']['kernels'])
print(d['extra']['configs'].get('C4 EvRep (3,440,640) f64, the pinned stand-in for the time surface (events_to_image.py:77-125)'))
"
tail -12 gpurun_out/${T}_traffic.csv
