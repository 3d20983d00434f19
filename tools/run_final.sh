set -x
mkdir -p gpurun_out
T=${TAG:-r02z}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/${T}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
tail -3 gpurun_out/${T}_bench_1gpu.err
# DRAM traffic of the headline step, caches NOT flushed between kernels
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:"k_route|k_sweep|k_sample_meta|k_tiled_setup|k_tiled_desc" -c 15 --csv --log-file gpurun_out/${T}_traffic.csv python tools/quick_bin.py --batch 256 --packed4 --methods tiled --steps 1 > gpurun_out/${T}_traffic.log 2>&1
# launch list of the bench command
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_route|k_sweep|k_sample_meta|k_tiled|k_stats|k_scatter|k_finalize|k_evrep|k_mask|k_gather|k_patch|k_ts_|k_view|k_block|k_swin|k_plane|k_norm|k_hot" -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_route|k_sweep" -c 2 -f -o gpurun_out/${T}_tiled python tools/quick_bin.py --batch 256 --packed4 --methods tiled --steps 1 > gpurun_out/${T}_ncu_tiled.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane$' -c 1 -f -o gpurun_out/${T}_plane python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 1 > gpurun_out/${T}_ncu_plane.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:"k_plane|k_sample_meta|k_tiled|k_route|k_sweep" -c 12 --csv --log-file gpurun_out/${T}_traffic_plane.csv python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 1 > gpurun_out/${T}_traffic_plane.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_evrep_sweep" -c 1 -f -o gpurun_out/${T}_evrep python tools/quick_evrep.py --steps 1 --only-tiled > gpurun_out/${T}_ncu_evrep.log 2>&1
cat gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_1gpu.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print(d['roofline']['frac'], d['roofline']['kernels'])
print(json.dumps(d['extra']['configs'],indent=0)[:5000])
"
cat gpurun_out/${T}_bench_ref.json | cut -c1-600
