set -x
mkdir -p gpurun_out
T=${TAG:-r02am}
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:"k_route|k_sweep" -c 6 --csv --log-file gpurun_out/${T}_traffic.csv python tools/quick_bin.py --batch 256 --packed4 --methods tiled --steps 1 > gpurun_out/${T}_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -c 1 -f -o gpurun_out/${T}_sweep python tools/quick_bin.py --batch 256 --packed4 --methods tiled --steps 1 > gpurun_out/${T}_ncu.log 2>&1
grep "k_sweep\|k_route" gpurun_out/${T}_traffic.csv | cut -d, -f5,13,15 | tail -8
