set -x
mkdir -p gpurun_out
T=r02br
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
for V in default ldg; do
  if [ $V != default ]; then cp build/variants/$V.so eventpretrain_b200/libeventpretrain_b200.so; fi
  echo "== $V" | tee -a gpurun_out/${T}_ab.log
  timeout 200 python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 20 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.log
  timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:'k_plane$' -c 2 --csv python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 1 2>&1 | grep "k_plane" | awk -F'","' '{print $13, $NF}' | tail -3 | tee -a gpurun_out/${T}_ab.log
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
