"""Small driver for ncu: a 32-sample slice of the bench workload, two passes of the binning path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import eventpretrain_b200 as ep  # noqa: E402

B = int(os.environ.get("EP_PROFILE_BATCH", "32"))
METHOD = os.environ.get("EP_PROFILE_METHOD", "auto")
ev = bench.make_batch_gpu(0, torch.device("cuda", 0), batch=B)
out = {}
for _ in range(2):
    out = ep.bin_events(ev, (bench.H, bench.W), num_bins=bench.BINS, voxel_sum=True, out=out, method=METHOD)
torch.cuda.synchronize()
print("events", ev.num_events, "checksum", float(out["voxel_sum"].sum()))
