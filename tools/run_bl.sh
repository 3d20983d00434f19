set -x
mkdir -p gpurun_out
T=r02zc
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_route|k_sweep|k_sample_meta|k_tiled|k_stats|k_scatter|k_finalize|k_evrep|k_mask|k_gather|k_patch|k_ts_|k_view|k_block|k_swin|k_plane|k_norm|k_hot" -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1
tail -3 gpurun_out/${T}_launches.csv | cut -c1-200
