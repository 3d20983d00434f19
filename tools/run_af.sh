set -x
mkdir -p gpurun_out
T=${TAG:-r02av}
cp eventpretrain_b200/libeventpretrain_b200.so /tmp/lib_default.so
: > gpurun_out/${T}_ab.log
for V in "-DEP_ABL_ROUTE_ATOM" "-DEP_ABL_ROUTE_STAGE" "-DEP_ABL_ROUTE_ATOM -DEP_ABL_ROUTE_STAGE"; do
  echo "== route ablation $V (results wrong by construction; pass1 = route)" >> gpurun_out/${T}_ab.log
  EP_NVCC_EXTRA="$V" timeout 600 python -m eventpretrain_b200.build --force > /dev/null 2>&1
  timeout 120 python - >> gpurun_out/${T}_ab.log 2>&1 <<'P'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import bench
import eventpretrain_b200 as ep
from eventpretrain_b200 import _lib
dev = torch.device('cuda', 0)
L = ep.load_library()
ev = bench.make_batch_gpu(0, dev, batch=256)
host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu(), ev.p.cpu(), ev.offsets.cpu(), ev.offsets_host, ev.t_div)
p4 = host.packed(4).to(dev)
out = ep.bin_events(p4, (480, 640), num_bins=5, voxel_sum=True, method='tiled')
for _ in range(2): ep.bin_events(p4, (480, 640), num_bins=5, voxel_sum=True, method='tiled', out=out)
torch.cuda.synchronize()
L.ep_profile_enable(1)
for _ in range(5): ep.bin_events(p4, (480, 640), num_bins=5, voxel_sum=True, method='tiled', out=out)
prof = _lib.ProfileStats(); L.ep_profile_read(prof); L.ep_profile_enable(0)
print('route %.3f ms sweep %.3f ms' % (prof.ms[0] / 5, prof.ms[1] / 5))
P
done
cp /tmp/lib_default.so eventpretrain_b200/libeventpretrain_b200.so
cat gpurun_out/${T}_ab.log
