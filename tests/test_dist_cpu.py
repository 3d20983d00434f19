"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sample sharding and the statistics all-reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _table(x):
    """(C,4) fp64 table (count, sum, sum of squares, max) of a CPU tensor: what the GPU kernels hand to the all-reduce."""
    xd = x.double().transpose(0, 1).reshape(x.shape[1], -1)
    return torch.stack([torch.full((x.shape[1],), float(xd.shape[1]), dtype=torch.float64), xd.sum(1), (xd * xd).sum(1), xd.amax(1)], 1)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    import eventpretrain_b200 as ep
    from eventpretrain_b200 import dist as epd
    r, w, _ = epd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and dist.is_initialized()
    rng = np.random.default_rng(0)
    counts = [10, 0, 7, 3, 12, 5]
    samples = [np.stack([rng.integers(0, 64, n), rng.integers(0, 48, n), np.sort(rng.uniform(0, 1, n)),
                         rng.integers(0, 2, n)], 1).astype(np.float64) for n in counts]
    ev = ep.pack_events(samples, pin=False)
    sh = ev.shard(rank, world)
    lo, hi = epd.shard_range(len(counts), rank, world)
    assert sh.batch == hi - lo and sh.num_events == sum(counts[lo:hi])
    # statistics: each rank holds its shard of a (B,C,H,W) tensor; the reduced result must equal the global one
    g = torch.Generator().manual_seed(1)
    full = torch.randn(len(counts), 3, 8, 8, generator=g)
    stats = _table(full[lo:hi])
    stats, _ = epd.allreduce_statistics(stats)
    ref = _table(full)
    fin, fin_ref = epd.finalize_statistics(stats), epd.finalize_statistics(ref)
    ok = all(torch.allclose(fin[k], fin_ref[k], rtol=1e-12, atol=1e-12) for k in fin)
    finish, works = epd.allreduce_statistics(_table(full[lo:hi]), async_op=True)
    for wk in works:
        wk.wait()
    ok = ok and torch.allclose(finish(), ref, rtol=1e-12)
    q.put((rank, bool(ok), int(sh.num_events)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] + res[1][2] == 37


def test_balance_by_events():
    from eventpretrain_b200 import dist as epd
    counts = [100, 1, 1, 1, 50, 49, 2]
    parts = epd.balance_by_events(counts, 2)
    assert sorted(sum(parts, [])) == list(range(7))
    loads = [sum(counts[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 4
    assert epd.shard_range(256, 3, 8) == (96, 128) and epd.shard_range(10, 2, 3) == (6, 10)
