"""EvRep (events_to_EvRep, dataset/dataset_utils/events_to_image.py:77-125) on the routed shared-memory path that the 4 B/event
packed transport layout takes (csrc/ep_binning_tiled.cu: transposed route + per-tile counting sort and replay), against the
global counting-sort kernels on the canonical layout (bit for bit) and the CPU oracle (bit for bit, E_T included)."""
import numpy as np
import pytest
import torch

from test_gpu_tiled import dense_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ep(native_lib):
    import eventpretrain_b200 as ep
    return ep


def check(ep, ev, samples, H, W, n_oracle=3, t_div=1e6):
    from oracle import events as oe
    canon = ep.evrep(ev.to("cuda"), (H, W), check=True)
    p4 = ev.packed(4).to("cuda")
    for rep in range(2):                             # the workspace (look-back flags, counters) is reused
        tiled = ep.evrep(p4, (H, W), check=True)
        assert torch.equal(tiled, canon), rep
    for b in range(min(n_oracle, len(samples))):
        s = samples[b]
        if len(s) == 0:
            assert not tiled[b].any()
            continue
        ref = oe.evrep(s[:, 0].astype(np.int64), s[:, 1].astype(np.int64), np.round(s[:, 2] * t_div) / t_div, s[:, 3], (W, H))
        assert np.array_equal(tiled[b].cpu().numpy(), ref, equal_nan=True), b
    return tiled


@pytest.mark.parametrize("H,W", [(44, 64), (100, 131), (180, 240), (260, 346), (33, 50), (7, 9)])
def test_evrep_tiled_equals_canonical_and_oracle(ep, H, W):
    rng = np.random.default_rng(100 + H)
    counts = [30000, 0, 7, 65000, 1, 12345]
    ev, samples = dense_batch(ep, rng, counts, H, W, hot=600)
    check(ep, ev, samples, H, W, n_oracle=6)


def test_evrep_tiled_empty_tiles_and_sparse_columns(ep):
    """Events only in a few columns far apart: empty tiles between them (look-back skips them), a sample whose first columns
    are empty, single-pixel samples."""
    rng = np.random.default_rng(5)
    H, W = 120, 400

    def edit(xs, ys, ts, ps):
        xs[0][:] = rng.choice([3, 197, 198, 399], xs[0].shape[0])
        xs[1][:] = 399
        ys[1][:] = 119
        xs[2][:] = rng.choice([150, 151], xs[2].shape[0])
    ev, samples = dense_batch(ep, rng, [20000, 3000, 9000, 500], H, W, edit=edit)
    check(ep, ev, samples, H, W, n_oracle=4)


def test_evrep_tiled_unsorted_blocks_and_offset_base(ep):
    """Blocks of 256 events out of time order (stamps of a pixel then arrive unsorted: the per-pixel sort is exercised) and a
    large tick base (t_base carries it; the fp64 quotient (t_base + ticks) / t_div must be the reference's stamp)."""
    rng = np.random.default_rng(6)
    H, W = 90, 160
    ev, samples = dense_batch(ep, rng, [50000, 20000, 8192 * 3], H, W, block_shuffle=True, t0=1_700_000_000_000)
    check(ep, ev, samples, H, W)


def test_evrep_tiled_hot_group_outside_staging(ep):
    """More stamps on 32 neighbouring cells than a warp's staging buffer holds, and one pixel with thousands of events."""
    rng = np.random.default_rng(8)
    H, W = 64, 96

    def edit(xs, ys, ts, ps):
        n = xs[0].shape[0]
        xs[0][: n // 2] = 40
        ys[0][: n // 2] = rng.integers(10, 14, n // 2)
        xs[1][100:4100] = 95
        ys[1][100:4100] = 63
    ev, samples = dense_batch(ep, rng, [40000, 30000], H, W, edit=edit)
    check(ep, ev, samples, H, W)


def test_evrep_tiled_config_size(ep):
    """BASELINE configs[3]: DSEC-shaped 640x440, ~2 M events per sample, microsecond stamps: at its own size, against the
    canonical path for the batch and the oracle for one sample."""
    rng = np.random.default_rng(4003)
    H, W = 440, 640
    ev, samples = dense_batch(ep, rng, [2_000_000, 1_900_000, 2_100_000], H, W)
    check(ep, ev, samples, H, W, n_oracle=1)


def test_evrep_tiled_bad_events(ep):
    rng = np.random.default_rng(9)
    H, W = 50, 70

    def edit(xs, ys, ts, ps):
        xs[0][5] = 70          # x == W: IndexError in the reference
        ys[0][9] = 50
    ev, _ = dense_batch(ep, rng, [5000], H, W, edit=edit)
    with pytest.raises(ep.BadEventsError):
        ep.evrep(ev.packed(4).to("cuda"), (H, W), check=True)


def test_evrep_tiled_dense_tile_takes_several_windows(ep):
    """More events on one column tile than the shared-memory window holds (47 000 stamps): the tile is swept in several
    ranges of cells, each pixel still differenced against the previous non-empty one across the range boundary."""
    rng = np.random.default_rng(12)
    H, W = 7, 9                                    # 63 pixels, one tile
    ev, samples = dense_batch(ep, rng, [200_000, 3, 120_000], H, W)
    check(ep, ev, samples, H, W)
    H, W = 40, 30                                  # several tiles? no: 30 columns x 40 rows = 1200 cells, one tile, ~250 events per pixel
    ev, samples = dense_batch(ep, rng, [300_000], H, W)
    check(ep, ev, samples, H, W)


def test_evrep_tiled_pixel_beyond_the_window_is_reported(ep):
    rng = np.random.default_rng(13)
    H, W = 20, 20

    def edit(xs, ys, ts, ps):
        xs[0][:50_000] = 7
        ys[0][:50_000] = 11                        # 50 000 events on one pixel: more than the window holds
    ev, _ = dense_batch(ep, rng, [60_000], H, W, edit=edit)
    with pytest.raises(OverflowError):
        ep.evrep(ev.packed(4).to("cuda"), (H, W), check=True)
