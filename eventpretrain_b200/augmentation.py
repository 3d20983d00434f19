"""Host-side pieces of dataset/augmentation/events_augment.py that parameterise the kernels.

Both are O(1) / O(N) numpy on the host in the reference and stay so; what matters is that their
results (window indices, coordinate scale) are fed to the fused binning kernel with the same
numerics: `events_reshape`'s Python-double scale is passed as `scale=(sx, sy)` and applied in fp64
before truncation (SURVEY.md §7 "fp64 coordinate-scale trap").
"""
import numpy as np


def get_random_index(args, events, is_train, seed=None):
    """dataset/augmentation/events_augment.py:5-20 — window [start, start+fix_events_num), global np.random."""
    if seed is not None:
        np.random.seed(seed)
    fix = args.fix_events_num if is_train else args.val_fix_events_num
    n = events.shape[0]
    if n > fix:
        start = np.random.randint(0, n - fix)
        return start, start + fix
    return 0, n


def events_reshape(events, sensor_w, sensor_h, input_w, input_h):
    """dataset/augmentation/events_augment.py:22-26 — in place, returns events."""
    events[:, 0] *= (input_w / sensor_w)
    events[:, 1] *= (input_h / sensor_h)
    return events


def reshape_scale(sensor_w, sensor_h, input_w, input_h):
    """The (sx, sy) pair to pass as `scale=` to the binning operators instead of mutating the events."""
    return (input_w / sensor_w, input_h / sensor_h)
