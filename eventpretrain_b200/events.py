"""Ragged event batches on the device and the batched stage-1 operators.

The reference bins one sample at a time on the CPU inside Dataset.__getitem__.  The
fast path here keeps raw events as a ragged structure-of-arrays batch (x,y u16 | t i64/f64 | p u8 +
offsets), built once by the collate step in pinned memory, and bins the whole batch on the GPU.
"""
import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._runtime import lib, ptr, require_cuda, stream_ptr, workspace

_NP_TAG = {np.dtype(np.uint8): _lib.EP_U8, np.dtype(np.int8): _lib.EP_I8, np.dtype(np.uint16): _lib.EP_U16,
           np.dtype(np.int16): _lib.EP_I16, np.dtype(np.int32): _lib.EP_I32, np.dtype(np.int64): _lib.EP_I64,
           np.dtype(np.float32): _lib.EP_F32, np.dtype(np.float64): _lib.EP_F64}
_TORCH_TAG = {torch.uint8: _lib.EP_U8, torch.int8: _lib.EP_I8, torch.uint16: _lib.EP_U16, torch.int16: _lib.EP_I16,
              torch.int32: _lib.EP_I32, torch.int64: _lib.EP_I64, torch.float32: _lib.EP_F32,
              torch.float64: _lib.EP_F64, torch.uint32: _lib.EP_U32}


@dataclass
class RaggedEvents:
    """Structure-of-arrays batch of event streams; sample b owns [offsets[b], offsets[b+1])."""
    x: torch.Tensor
    y: torch.Tensor
    t: torch.Tensor
    p: torch.Tensor
    offsets: torch.Tensor          # int64 (B+1), same device as the data
    offsets_host: np.ndarray       # int64 (B+1)
    t_div: float = 1.0             # timestamp value = t / t_div (int64 microseconds, t_div=1e6 -> seconds)
    t_base: torch.Tensor = None    # transport layouts only: int64 (B,) per-sample base ticks.  compact(): t is uint32
                                   # (relative ticks | polarity << 31) and p is None.  packed(): x is one uint32 word
                                   # per event, y is None, t the low tick byte, p the per-1024-block tick offsets

    @property
    def batch(self):
        return int(self.offsets_host.shape[0] - 1)

    @property
    def num_events(self):
        return int(self.offsets_host[-1] - self.offsets_host[0])

    @property
    def device(self):
        return self.x.device

    def _arrays(self):
        return [a for a in (self.x, self.y, self.t, self.p, self.offsets, self.t_base) if a is not None]

    def nbytes(self):
        return sum(int(a.numel()) * a.element_size() for a in self._arrays())

    def _map(self, f):
        g = lambda a: None if a is None else f(a)
        return RaggedEvents(g(self.x), g(self.y), g(self.t), g(self.p), g(self.offsets), self.offsets_host, self.t_div,
                            g(self.t_base))

    def to(self, device, non_blocking=True):
        return self._map(lambda a: a.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda a: a.pin_memory())

    def compact(self, native=True, threads=0):
        """Host-side repack into the 8 B/event transport layout: u16 x,y + u32 (ticks relative to the sample's first
        stamp | polarity << 31).  Needs int64 tick stamps, p in {0,1} and < 2^31 ticks between a sample's smallest and
        largest stamp.  Binning results are bit-identical to the int64 layout; H2D traffic drops from 13 to 8 B/event.
        Host-resident batches go through the library's threaded `ep_pack_transport_host`; native=False is the numpy
        statement of the same rule."""
        if self.t_base is not None:
            return self
        if self.t.dtype != torch.int64 or self.p.dtype != torch.uint8 or self.x.dtype != torch.uint16:
            raise TypeError("compact() needs the canonical u16 / int64-tick / u8 layout")
        off = self.offsets_host
        if native and not self.x.is_cuda:
            pin = self.x.is_pinned() and torch.cuda.is_available()
            rel = torch.empty(self.t.shape[0], dtype=torch.uint32, pin_memory=pin)
            base = torch.empty(self.batch, dtype=torch.int64, pin_memory=pin)
            o = np.ascontiguousarray(off, np.int64)
            t, p = self.t.contiguous(), self.p.contiguous()
            rc = _lib.load().ep_pack_transport_host(None, None, t.data_ptr(), p.data_ptr(), o.ctypes.data, self.batch, 8,
                                                    rel.data_ptr(), None, None, base.data_ptr(), int(threads))
            if rc == _lib.EP_EUNSUPPORTED:
                raise ValueError("compact() needs p in {0,1} and fewer than 2^31 ticks between a sample's smallest and largest stamp")
            _lib.check(rc, "ep_pack_transport_host")
            return RaggedEvents(self.x, self.y, rel, None, self.offsets, off, self.t_div, base)
        t = self.t.cpu().numpy()
        p = self.p.cpu().numpy()
        B = self.batch
        base = np.zeros(B, np.int64)
        rel = np.empty(t.shape[0], np.uint32)
        for b in range(B):
            lo, hi = int(off[b]), int(off[b + 1])
            if hi > lo:
                base[b] = t[lo:hi].min()
                d = t[lo:hi] - base[b]
                if d.max() >= (1 << 31):
                    raise ValueError("sample spans more than 2^31 ticks")
                if p[lo:hi].max() > 1:
                    raise ValueError("compact() needs polarity in {0, 1}")
                rel[lo:hi] = d.astype(np.uint32)
        rel |= p.astype(np.uint32) << np.uint32(31)
        mk = (lambda a: a.pin_memory()) if (self.x.is_pinned() or self.x.is_cuda) and torch.cuda.is_available() else (lambda a: a)
        return RaggedEvents(self.x, self.y, mk(torch.from_numpy(rel)).to(self.x.device), None, self.offsets, off, self.t_div,
                            mk(torch.from_numpy(base)).to(self.x.device))

    def take(self, b0, b1):
        """Standalone batch of the samples [b0, b1): copies of their array ranges, offsets rebased to 0 (canonical /
        generic / compact layouts; a packed() batch cannot be cut, its block offsets are tied to array positions)."""
        if self.t_base is not None and self.y is None:
            raise TypeError("take() on a packed() batch: take first, pack afterwards")
        lo, hi = int(self.offsets_host[b0]), int(self.offsets_host[b1])
        cut = lambda a: None if a is None else a[lo:hi].clone()
        off = (self.offsets_host[b0:b1 + 1] - lo).astype(np.int64)
        return RaggedEvents(cut(self.x), cut(self.y), cut(self.t), cut(self.p), torch.from_numpy(off.copy()).to(self.offsets.device),
                            off, self.t_div, None if self.t_base is None else self.t_base[b0:b1].clone())

    def pack_block(self):
        """Events per tick-offset block of a packed() batch: 1024 (5 B/event form) or 256 (4 B/event form)."""
        return 1024 if self.t is not None else 256

    def transport(self, threads=0):
        """Densest lossless transport layout this batch fits: packed 4 B/event, else packed 5 B/event, else compact 8 B/event."""
        for make in (lambda: self.packed(4, threads=threads), lambda: self.packed(5, threads=threads), lambda: self.compact(threads=threads)):
            try:
                return make()
            except ValueError:
                continue
        return self

    def packed(self, nbytes=5, native=True, threads=0):
        """Host-side repack into a packed transport layout (include/eventpretrain_b200.h, ep_events_soa.t_base): one
        uint32 x | y << 11 | polarity << 22 | tick bits << 23 per event; `ticks` count from the sample's base inside the
        block of the arrays where the sample starts and from base + a per-block offset afterwards.
          nbytes=5: blocks of 1024 events, 17-bit ticks = word bits | one extra byte per event
          nbytes=4: blocks of 256 events, 9-bit ticks in the word (dense streams)
        Needs int64 tick stamps, p in {0,1}, x,y < 2048 and a block's stamps (of one sample) within the tick range (raises
        ValueError otherwise: use the next wider layout, see transport()).  Binning results are bit-identical to the int64
        layout; H2D traffic drops from 13 to 5 or 4 B/event.  Host-resident batches are packed by the library's threaded
        `ep_pack_transport_host` (native=True; `threads` <= 0: all cores) straight into pinned buffers; native=False is the
        numpy statement of the same rule (what the tests compare it with)."""
        if nbytes not in (4, 5):
            raise ValueError("packed layouts have 4 or 5 bytes per event")
        if self.t_base is not None and self.y is None:
            if (self.t is None) != (nbytes == 4):
                raise TypeError("already packed with the other width")
            return self
        if self.t_base is not None or self.t.dtype != torch.int64 or self.p.dtype != torch.uint8 or self.x.dtype != torch.uint16:
            raise TypeError("packed() needs the canonical u16 / int64-tick / u8 layout")
        if int(self.offsets_host[0]) != 0:
            raise ValueError("packed() needs a whole batch (offsets[0] == 0): the block offsets are tied to array positions")
        off = self.offsets_host
        if native and not self.x.is_cuda:
            return self._packed_native(nbytes, threads)
        x = self.x.cpu().numpy().astype(np.uint32)
        y = self.y.cpu().numpy().astype(np.uint32)
        t = self.t.cpu().numpy()
        p = self.p.cpu().numpy().astype(np.uint32)
        n = t.shape[0]
        if n and (x.max() >= 2048 or y.max() >= 2048):
            raise ValueError("packed() needs x, y < 2048")
        if n and p.max() > 1:
            raise ValueError("packed() needs polarity in {0, 1}")
        B, K = self.batch, (1024 if nbytes == 5 else 256)
        tick_bits = 17 if nbytes == 5 else 9
        counts = np.diff(off)
        base = np.zeros(B, np.int64)
        nz = counts > 0
        if n:
            base[nz] = np.minimum.reduceat(t, off[:-1][nz])     # per-sample smallest stamp (empty samples skipped)
        sid = np.repeat(np.arange(B), counts)                    # owning sample of every event
        rel = t - base[sid]
        g = np.arange(n) // K                                    # 1024-block of every event
        first_blk = (off[:-1] // K)[sid]                         # block where the event's sample starts
        blk = np.zeros((n + K - 1) // K, np.uint32)
        if n:
            starts = np.arange(0, n, K)
            own = sid == sid[starts][g]                          # event belongs to the sample owning its block's first slot
            m = np.minimum.reduceat(np.where(own, rel, np.iinfo(np.int64).max), starts)
            if m.max(initial=0) >= (1 << 32) and (m < np.iinfo(np.int64).max).any() and m[m < np.iinfo(np.int64).max].max() >= (1 << 32):
                raise ValueError("sample spans more than 2^32 ticks")
            blk = np.where(m < (1 << 32), m, 0).astype(np.uint32)
        ticks = rel - np.where(g == first_blk, 0, blk[g].astype(np.int64)) if n else rel
        if n and ticks.min() < 0:
            raise ValueError("stamps too far out of order for the packed layout")
        if n and ticks.max() >= (1 << tick_bits):
            raise ValueError(f"a {K}-event block spans 2^{tick_bits} ticks or more")
        tk = ticks.astype(np.uint32)
        dev = self.x.device
        mk = (lambda a: a.pin_memory()) if (self.x.is_pinned() or self.x.is_cuda) and torch.cuda.is_available() else (lambda a: a)
        cv = lambda a: mk(torch.from_numpy(np.ascontiguousarray(a))).to(dev)
        if nbytes == 5:
            w = x | (y << np.uint32(11)) | (p << np.uint32(22)) | ((tk >> np.uint32(8)) << np.uint32(23))
            tl = cv((tk & np.uint32(0xff)).astype(np.uint8))
        else:
            w = x | (y << np.uint32(11)) | (p << np.uint32(22)) | (tk << np.uint32(23))
            tl = None
        return RaggedEvents(cv(w), None, tl, cv(blk), self.offsets, off, self.t_div, cv(base))

    def _packed_native(self, nbytes, threads):
        n, B, K = int(self.offsets_host[-1]), self.batch, (1024 if nbytes == 5 else 256)
        pin = self.x.is_pinned() and torch.cuda.is_available()
        mk = lambda size, dt: torch.empty(size, dtype=dt, pin_memory=pin)
        w, blk, base = mk(n, torch.uint32), mk((n + K - 1) // K, torch.uint32), mk(B, torch.int64)
        tl = mk(n, torch.uint8) if nbytes == 5 else None
        off = np.ascontiguousarray(self.offsets_host, np.int64)
        x, y, t, p = (a.contiguous() for a in (self.x, self.y, self.t, self.p))
        rc = _lib.load().ep_pack_transport_host(x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(), off.ctypes.data, B, nbytes,
                                                w.data_ptr(), ptr(tl), blk.data_ptr(), base.data_ptr(), int(threads))
        if rc == _lib.EP_EUNSUPPORTED:
            raise ValueError(f"batch does not fit the {nbytes} B/event packed layout (x, y < 2048, p in {{0,1}}, "
                             f"{K}-event blocks within 2^{17 if nbytes == 5 else 9} ticks)")
        _lib.check(rc, "ep_pack_transport_host")
        return RaggedEvents(w, None, tl, blk, self.offsets, self.offsets_host, self.t_div, base)

    def unpack_host(self):
        """Decode a packed() batch back to (x, y, t_ticks, p) numpy arrays (tests / debugging)."""
        if self.t_base is None or self.y is not None:
            raise TypeError("not a packed() batch")
        w = self.x.cpu().numpy().astype(np.uint32)
        blk = self.p.cpu().numpy().astype(np.int64)
        base = self.t_base.cpu().numpy()
        off, K = self.offsets_host, self.pack_block()
        ticks = (w >> np.uint32(23)).astype(np.int64)
        if self.t is not None:
            ticks = (ticks << 8) | self.t.cpu().numpy().astype(np.int64)
        t = np.zeros(w.shape[0], np.int64)
        for b in range(self.batch):
            lo, hi = int(off[b]), int(off[b + 1])
            if hi <= lo:
                continue
            g = np.arange(lo, hi) // K
            add = np.where(g == lo // K, 0, blk[g])
            t[lo:hi] = base[b] + add + ticks[lo:hi]
        return (w & np.uint32(0x7ff)).astype(np.uint16), ((w >> np.uint32(11)) & np.uint32(0x7ff)).astype(np.uint16), t, \
            ((w >> np.uint32(22)) & np.uint32(1)).astype(np.uint8)

    def shard(self, rank, world_size):
        """Samples [rank*B/G, (rank+1)*B/G): DistributedSampler-style contiguous split (main_pretrain.py:218-220).
        Slices are views; offsets keep their absolute values (ep_events_soa allows offsets[0] > 0)."""
        B = self.batch
        lo, hi = (B * rank) // world_size, (B * (rank + 1)) // world_size
        return RaggedEvents(self.x, self.y, self.t, self.p, self.offsets[lo:hi + 1], self.offsets_host[lo:hi + 1],
                            self.t_div, None if self.t_base is None else self.t_base[lo:hi])

    def _desc(self):
        d = _lib.EventsSoa()
        d.x, d.y, d.t, d.p = ptr(self.x), ptr(self.y), ptr(self.t), ptr(self.p)
        if self.y is not None and self.x.dtype != self.y.dtype:
            raise TypeError("x and y must share a dtype")
        d.xy_dtype = _TORCH_TAG[self.x.dtype]
        d.t_dtype = _TORCH_TAG[self.t.dtype] if self.t is not None else 0
        d.p_dtype = _TORCH_TAG[self.p.dtype] if self.p is not None else 0
        d.t_base = ptr(self.t_base)
        d.batch = self.batch
        d.t_div = float(self.t_div)
        d.offsets = ptr(self.offsets)
        self._off_host = np.ascontiguousarray(self.offsets_host, np.int64)
        d.offsets_host = self._off_host.ctypes.data
        return d


def pack_events(samples, t_div=1.0, canonical=True, pin=True):
    """Collate per-sample (N,4) x,y,t,p arrays (the reference's event format) into a host-side ragged SoA batch.

    canonical=True stores x,y as uint16 and p as uint8 (requires integer coordinates in [0, 65535] and
    p in {0,1}); timestamps keep their dtype (float64, or int64 ticks with `t_div`).  canonical=False
    keeps float coordinates (events with sub-pixel jitter from erase_and_add_events).
    """
    counts = np.array([len(s) for s in samples], np.int64)
    offsets = np.zeros(len(samples) + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    ev = np.concatenate([np.asarray(s) for s in samples], axis=0) if len(samples) else np.zeros((0, 4))
    if canonical:
        x = ev[:, 0].astype(np.uint16)
        y = ev[:, 1].astype(np.uint16)
        if not (np.array_equal(x, ev[:, 0]) and np.array_equal(y, ev[:, 1])):
            raise ValueError("canonical packing needs integer coordinates in [0, 65535]")
        if not np.isin(ev[:, 3], (0, 1)).all():
            raise ValueError("canonical packing needs polarity in {0, 1}")
        p = ev[:, 3].astype(np.uint8)
    else:
        x = np.ascontiguousarray(ev[:, 0])
        y = np.ascontiguousarray(ev[:, 1])
        p = np.ascontiguousarray(ev[:, 3])
    t = np.ascontiguousarray(ev[:, 2])
    return from_soa(x, y, t, p, offsets, t_div=t_div, pin=pin)


def collate_events(samples, ticks_per_unit=1.0e6, pin=True, threads=0):
    """The collate step as native threaded code (`ep_collate_aos_host`): per-sample (N,4) x,y,t,p arrays (float64 or float32,
    the reference's event format) -> canonical ragged batch with uint16 coordinates, uint8 polarity and int64 tick stamps
    `rint(t * ticks_per_unit)` (t_div = ticks_per_unit, so stamps in seconds at the default become microsecond ticks),
    written straight into pinned buffers.  Raises ValueError for batches the canonical layout cannot hold (fractional or
    out-of-range coordinates, polarity outside {0,1}): those go through pack_events(canonical=False)."""
    B = len(samples)
    if B == 0:
        raise ValueError("empty batch")
    # float32 rows only when every sample is float32 (DDD17 / DVS128-Gesture); one float64 sample promotes the batch, so that
    # stamps in seconds never lose their microseconds to a narrower neighbour's dtype
    dt = np.dtype(np.float32) if all(np.asarray(s).dtype == np.float32 for s in samples) else np.dtype(np.float64)
    arrs = [np.ascontiguousarray(s, dt).reshape(-1, 4) for s in samples]
    counts = np.array([a.shape[0] for a in arrs], np.int64)
    n = int(counts.sum())
    pin = pin and torch.cuda.is_available()
    mk = lambda size, d: torch.empty(size, dtype=d, pin_memory=pin)
    x, y, t, p, off = mk(n, torch.uint16), mk(n, torch.uint16), mk(n, torch.int64), mk(n, torch.uint8), mk(B + 1, torch.int64)
    ptrs = (ctypes.c_void_p * B)(*[a.ctypes.data for a in arrs])
    rc = _lib.load().ep_collate_aos_host(ptrs, counts.ctypes.data, B, _lib.EP_F64 if dt == np.float64 else _lib.EP_F32,
                                         float(ticks_per_unit), x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(),
                                         off.data_ptr(), int(threads))
    if rc == _lib.EP_EUNSUPPORTED:
        raise ValueError("canonical collate needs integer coordinates in [0, 65535] and polarity in {0, 1}")
    _lib.check(rc, "ep_collate_aos_host")
    return RaggedEvents(x, y, t, p, off, offsets_host=off.numpy().copy(), t_div=float(ticks_per_unit))


def collate_transport(samples, ticks_per_unit=1.0e6, pin=True, threads=0):
    """collate_events(...).transport() for the common case in one pass over the rows (`ep_collate_transport4_host`): per-sample
    (N,4) x,y,t,p arrays -> the 4 B/event packed layout, reading the 32-byte rows once instead of writing and re-reading the
    13 B/event canonical arrays in between.  Covers time-sorted samples that fit the 4 B layout (x, y < 2048, p in {0,1},
    256-event blocks within 512 ticks) and produces the same arrays bit for bit; anything else falls back to the two-step
    form, which picks the next wider layout."""
    B = len(samples)
    if B == 0:
        raise ValueError("empty batch")
    dt = np.dtype(np.float32) if all(np.asarray(s).dtype == np.float32 for s in samples) else np.dtype(np.float64)
    arrs = [np.ascontiguousarray(s, dt).reshape(-1, 4) for s in samples]
    counts = np.array([a.shape[0] for a in arrs], np.int64)
    n = int(counts.sum())
    pinned = pin and torch.cuda.is_available()
    mk = lambda size, d: torch.empty(size, dtype=d, pin_memory=pinned)
    w, blk, base, off = mk(n, torch.uint32), mk((n + 255) // 256, torch.uint32), mk(B, torch.int64), mk(B + 1, torch.int64)
    ptrs = (ctypes.c_void_p * B)(*[a.ctypes.data for a in arrs])
    rc = _lib.load().ep_collate_transport4_host(ptrs, counts.ctypes.data, B, _lib.EP_F64 if dt == np.float64 else _lib.EP_F32,
                                                float(ticks_per_unit), w.data_ptr(), blk.data_ptr(), base.data_ptr(), off.data_ptr(),
                                                int(threads))
    if rc == _lib.EP_EUNSUPPORTED:
        return collate_events(arrs, ticks_per_unit, pin, threads).transport(threads=threads)
    _lib.check(rc, "ep_collate_transport4_host")
    return RaggedEvents(w, None, None, blk, off, off.numpy().copy(), float(ticks_per_unit), base)


class EventCollator:
    """`collate_fn` for a torch DataLoader whose dataset returns the raw event window instead of binning it
    (the reference bins per sample on the CPU inside `__getitem__`, e.g.
    pr_n_imagenet_dataset.py:85-87).  Each item is a dict; `events_key` holds the reference-format (N,4) x,y,t,p array.
    The events of the batch become one RaggedEvents (native threaded collate; the densest transport layout that fits when
    layout="transport", the canonical 13 B/event layout when layout="canonical"); batches the canonical layout cannot
    hold (fractional coordinates after erase_and_add_events, polarity -1) fall back to the generic layout.  Every other key
    goes through torch's default_collate.  Workers return unpinned buffers; `DataLoader(pin_memory=True)` pins the batch
    through `RaggedEvents.pin_memory()`.  threads=1 suits many worker processes; 0 uses every core of one."""

    def __init__(self, events_key="events", ticks_per_unit=1.0e6, layout="transport", threads=1):
        if layout not in ("transport", "canonical"):
            raise ValueError("layout is 'transport' or 'canonical'")
        self.events_key, self.ticks_per_unit, self.layout, self.threads = events_key, float(ticks_per_unit), layout, int(threads)

    def __call__(self, batch):
        from torch.utils.data import default_collate
        key = self.events_key
        samples = [item[key] for item in batch]
        rest = [{k: v for k, v in item.items() if k != key} for item in batch]
        out = default_collate(rest) if rest and rest[0] else {}
        try:
            if self.layout == "transport":
                ev = collate_transport(samples, self.ticks_per_unit, pin=False, threads=self.threads)
            else:
                ev = collate_events(samples, self.ticks_per_unit, pin=False, threads=self.threads)
        except ValueError:
            ev = pack_events(samples, canonical=False, pin=False)
        out[key] = ev
        return out


def from_soa(x, y, t, p, offsets, t_div=1.0, pin=True):
    """Wrap host SoA numpy arrays (any dtype tag of ep_dtype) as a host-resident RaggedEvents."""
    off = np.ascontiguousarray(offsets, np.int64)
    tens = [torch.from_numpy(np.ascontiguousarray(a)) for a in (x, y, t, p)]
    tens.append(torch.from_numpy(off.copy()))
    if pin and torch.cuda.is_available():
        tens = [a.pin_memory() for a in tens]
    return RaggedEvents(*tens, offsets_host=off, t_div=t_div)


_METHOD = {None: 0, "auto": 0, "global": _lib.EP_BIN_FORCE_GLOBAL, "tiled": _lib.EP_BIN_FORCE_TILED,
           "plane": _lib.EP_BIN_FORCE_PLANE}


def _bin_params(size, num_bins, count_channels, scale, time_f32, method=None):
    prm = _lib.BinParams()
    prm.height, prm.width = int(size[0]), int(size[1])
    prm.num_bins, prm.count_channels = int(num_bins), int(count_channels)
    prm.scale_x, prm.scale_y = float(scale[0]), float(scale[1])
    prm.time_f32 = int(bool(time_f32))
    prm.flags = _METHOD[method]
    return prm


class BadEventsError(IndexError):
    """Out-of-range coordinates / unsupported polarity (the reference raises IndexError / RuntimeError)."""


def _raise_bad(bad):
    v = int(bad.item())
    if v & 0x80000000:
        raise OverflowError("voxel accumulator headroom exceeded (>= 2^18 net same-polarity events on one pixel "
                            "within one temporal interval)")
    if v:
        raise BadEventsError(f"{v} event(s) index outside the grid or carry a polarity outside {{-1,0,1}}")


def bin_events(ev, size, num_bins=0, count_channels=0, scale=(1.0, 1.0), voxel_sum=False, time_f32=False,
               check=False, out=None, method=None, stats=False, bad_out=None):
    """Batched events -> tensors: voxel grid (B,num_bins,H,W), optional voxel.sum(0) plane (B,1,H,W) and/or
    polarity count frame (B,count_channels,H,W); one C-ABI call (ep_bin_events).

    Returns a dict with the keys that were requested: 'voxel', 'voxel_sum', 'count'.
    stats=True adds 'stats': (num_bins + 1, 4) fp64 rows (element count, sum, sum of squares, max) per voxel channel and, in
    the last row, of the sum plane (valid when voxel_sum) — the table dist.allreduce_statistics reduces over the ranks; on
    the tiled path it is a by-product of the kernel that writes the planes.
    check=True synchronises and raises for events the reference would have raised on.
    method: None/"auto" (for the 4 B/event packed layout without count frames: the whole-plane shared-memory kernels when the
    grid's plane fits one SM — 224 x 224, 240 x 180 — else route + sweep; the global-RED kernels for everything else),
    "global", "tiled", "plane" — same results bit for bit.
    bad_out: optional zeroed int32 CUDA tensor that receives the number of events the reference would have raised on
    (bit 31: accumulator overflow), without a synchronisation.
    """
    require_cuda(ev.x)
    dev = ev.device
    B, (H, W) = ev.batch, size
    prm = _bin_params(size, num_bins, count_channels, scale, time_f32, method)
    out = {} if out is None else out
    if num_bins and "voxel" not in out:
        out["voxel"] = torch.empty((B, num_bins, H, W), dtype=torch.float32, device=dev)
    if voxel_sum and "voxel_sum" not in out:
        out["voxel_sum"] = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    if count_channels and "count" not in out:
        out["count"] = torch.empty((B, count_channels, H, W), dtype=torch.float32, device=dev)
    L = lib()
    desc = ev._desc()
    nbytes = L.ep_bin_events_workspace_bytes_for(ctypes.byref(desc), ctypes.byref(prm))
    ws = workspace(nbytes, dev, "bin")
    bad = bad_out if bad_out is not None else (torch.zeros(1, dtype=torch.int32, device=dev) if check else None)
    if stats:
        if not num_bins:
            raise ValueError("stats=True needs a voxel grid")
        if "stats" not in out:
            out["stats"] = torch.zeros((num_bins + 1, 4), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = L.ep_bin_events_stats(stream_ptr(dev), ctypes.byref(desc), ctypes.byref(prm), ptr(out.get("voxel")),
                                   ptr(out.get("voxel_sum")) if voxel_sum else 0, ptr(out.get("count")), ws.data_ptr(),
                                   ws.numel(), ptr(bad), ptr(out.get("stats")) if stats else 0)
    _lib.check(rc, "ep_bin_events")
    if check:
        _raise_bad(bad)
    return out


def bin_events_aos(events, size, num_bins=0, count_channels=0, scale=(1.0, 1.0), voxel_sum=False, check=True):
    """Single (N,4) x,y,t,p CUDA tensor, fp64 or fp32 (fp32 => the reference's fp32 time arithmetic)."""
    require_cuda(events)
    if events.dim() != 2 or events.shape[1] != 4:
        raise AssertionError("events must be (N, 4)")   # events_to_voxel_grid.py:9
    if events.shape[0] == 0:
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")   # events[0, 2] in the reference
    events = events.contiguous()
    dev = events.device
    H, W = size
    prm = _bin_params(size, num_bins, count_channels, scale, events.dtype == torch.float32)
    d = _lib.EventsAos()
    d.events, d.dtype, d.n = events.data_ptr(), _TORCH_TAG[events.dtype], events.shape[0]
    out = {}
    if num_bins:
        out["voxel"] = torch.empty((num_bins, H, W), dtype=torch.float32, device=dev)
    if voxel_sum:
        out["voxel_sum"] = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    if count_channels:
        out["count"] = torch.empty((count_channels, H, W), dtype=torch.float32, device=dev)
    L = lib()
    nbytes = L.ep_bin_events_workspace_bytes(ctypes.byref(prm), 1, None)
    ws = workspace(nbytes, dev, "bin")
    bad = torch.zeros(1, dtype=torch.int32, device=dev) if check else None
    with torch.cuda.device(dev):
        rc = L.ep_bin_events_aos(stream_ptr(dev), ctypes.byref(d), ctypes.byref(prm), ptr(out.get("voxel")),
                                 ptr(out.get("voxel_sum")), ptr(out.get("count")), ws.data_ptr(), ws.numel(), ptr(bad))
    _lib.check(rc, "ep_bin_events_aos")
    if check:
        _raise_bad(bad)
    return out


def normalise(img, mode):
    """In-place per-sample normaliser on (B,C,H,W) or (C,H,W): mode 'count' | 'mem' | 'mem_guard'."""
    require_cuda(img)
    code = {"count": _lib.EP_NORM_COUNT, "mem": _lib.EP_NORM_MEM, "mem_guard": _lib.EP_NORM_MEM_GUARD}[mode]
    v = img if img.dim() == 4 else img.unsqueeze(0)
    if not v.is_contiguous() or v.dtype != torch.float32:
        raise ValueError("normalise works in place on a contiguous float32 tensor")
    B, C, H, W = v.shape
    L = lib()
    ws = workspace(L.ep_normalise_workspace_bytes(B, C), v.device, "norm")
    with torch.cuda.device(v.device):
        rc = L.ep_normalise(stream_ptr(v.device), v.data_ptr(), B, C, H, W, code, ws.data_ptr(), ws.numel())
    _lib.check(rc, "ep_normalise")
    return img


def mem_hotpixel(hist, num_stds=10.0, divide_by=1.0):
    """In-place remove_hot_pixel_mem on (B,3,H,W) or (3,H,W), optionally fusing the callers' `/ 255`."""
    require_cuda(hist)
    v = hist if hist.dim() == 4 else hist.unsqueeze(0)
    if not v.is_contiguous() or v.dtype != torch.float32 or v.shape[1] != 3:
        raise ValueError("mem_hotpixel works in place on a contiguous float32 (B,3,H,W) tensor")
    B, _, H, W = v.shape
    L = lib()
    ws = workspace(L.ep_mem_hotpixel_workspace_bytes(B), v.device, "hot")
    with torch.cuda.device(v.device):
        rc = L.ep_mem_hotpixel(stream_ptr(v.device), v.data_ptr(), B, H, W, float(divide_by), float(num_stds),
                               ws.data_ptr(), ws.numel())
    _lib.check(rc, "ep_mem_hotpixel")
    return hist


def evrep(ev, size, check=False, out=None):
    """Batched EvRep: (B,3,H,W) float64 = [E_C, E_I, E_T] (events_to_image.py:77-125).  Canonical / generic layouts take the
    global counting sort; the 4 B packed transport layout (`RaggedEvents.transport()` on dense streams) takes the routed
    shared-memory path, bit-identical (stamp value = (t_base + ticks) / t_div in both)."""
    require_cuda(ev.x)
    dev = ev.device
    H, W = size
    B = ev.batch
    if out is None:
        out = torch.empty((B, 3, H, W), dtype=torch.float64, device=dev)
    L = lib()
    desc = ev._desc()
    ws = workspace(L.ep_evrep_workspace_bytes_for(ctypes.byref(desc), H, W), dev, "evrep")
    bad = torch.zeros(1, dtype=torch.int32, device=dev) if check else None
    with torch.cuda.device(dev):
        rc = L.ep_evrep(stream_ptr(dev), ctypes.byref(desc), H, W, out.data_ptr(), ws.data_ptr(), ws.numel(), ptr(bad))
    _lib.check(rc, "ep_evrep")
    if check:
        v = int(bad.item())
        if v & 0x80000000:
            raise OverflowError("EvRep: more than 65535 events on one pixel of one sample (routed path)")
        if v & 0x40000000:
            raise BadEventsError("EvRep: a stamp lies before, or 2^32 ticks or more after, its sample's first row (routed path)")
        if v & 0x20000000:
            raise OverflowError("EvRep: more than 47000 events on one pixel, or more than 752k events on one column tile (routed path)")
        _raise_bad(bad)
    return out


def time_surface(ev, size, tau, t_ref=None, check=False):
    """(B,2,H,W) f32 exponential time surface, exp(-(t_ref - t_last)/tau) per polarity and pixel (0 where no event).
    No counterpart exists in the reference (parity unpinned; self-oracle oracle/stage3_np.py:time_surface)."""
    require_cuda(ev.x)
    dev = ev.device
    H, W = size
    B = ev.batch
    out = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
    L = lib()
    ws = workspace(L.ep_time_surface_workspace_bytes(B, H, W), dev, "tsurf")
    bad = torch.zeros(1, dtype=torch.int32, device=dev) if check else None
    tr = None if t_ref is None else torch.as_tensor(t_ref, dtype=torch.float64, device=dev).contiguous()
    desc = ev._desc()
    with torch.cuda.device(dev):
        rc = L.ep_time_surface(stream_ptr(dev), ctypes.byref(desc), H, W, float(tau), ptr(tr), out.data_ptr(), ws.data_ptr(),
                               ws.numel(), ptr(bad))
    _lib.check(rc, "ep_time_surface")
    if check:
        _raise_bad(bad)
    return out
