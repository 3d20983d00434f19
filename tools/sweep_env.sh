#!/bin/bash
for cfg in "0 64" "1 64" "1 40" "1 76"; do
  set -- $cfg
  EP_PERSIST_L2=$1 EP_L2_GROUP_MB=$2 python bench.py --steps 6 --warmup 3 --method global 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('global persist=$1 group_mb=$2', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernels'].items()})"
done
