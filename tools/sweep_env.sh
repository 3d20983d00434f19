#!/bin/bash
for mb in 48 80 128; do
  EP_L2_GROUP_MB=$mb python bench.py --steps 6 --warmup 3 --method global 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('global fused L2_GROUP_MB=$mb', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernels'].items()}, d['extra'])"
done
