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


def erase_and_add_events(args, events, size=None):
    """dataset/augmentation/events_augment.py:28-55 — drop 0.1-1 % of the events, add 0.1-1 % jittered copies
    (N(0,1.5 px), N(0,1 ms)), clip to the sensor, re-sort by time.  Host numpy on purpose: the draws come from the
    global numpy RNG in the reference's order, so a seeded run selects the same events; the result has fractional
    coordinates, which the binning kernels truncate like the reference (generic SoA / AoS loaders)."""
    sensor_h, sensor_w = size[0], size[1]
    n = events.shape[0]
    lo, hi = int(0.001 * n), int(0.01 * n)
    if hi <= 0:
        return events
    erase_num = np.random.randint(lo, hi)
    erase_index = np.sort(np.random.choice(np.arange(n), size=erase_num, replace=False))
    add_num = np.random.randint(lo, hi)
    jittered = np.concatenate((events[:, [0]] + np.random.normal(0, 1.5, size=(n, 1)),
                               events[:, [1]] + np.random.normal(0, 1.5, size=(n, 1)),
                               events[:, [2]] + np.random.normal(0, 0.001, size=(n, 1)),
                               events[:, [3]]), 1)
    add = jittered[np.random.choice(np.arange(n), size=add_num, replace=False)]
    add[:, 0] = np.clip(add[:, 0], 0, sensor_w - 1)
    add[:, 1] = np.clip(add[:, 1], 0, sensor_h - 1)
    out = np.concatenate((np.delete(events, erase_index, axis=0), add))
    return out[out[:, 2].argsort()]


def add_noise_events(args, events, size):
    """dataset/augmentation/events_augment.py:57-77 — add 10-50 % uniform noise events inside the time window."""
    sensor_h, sensor_w = size[0], size[1]
    n = events.shape[0]
    add_num = np.random.randint(int(0.1 * n), int(0.5 * n))
    noise = np.concatenate((np.random.randint(0, sensor_w, size=(n, 1)), np.random.randint(0, sensor_h, size=(n, 1)),
                            np.random.uniform(events[0, 2], events[-1, 2], size=(n, 1)),
                            np.random.randint(0, 2, size=(n, 1))), 1)
    add = noise[np.random.choice(np.arange(n), size=add_num, replace=False)]
    add[:, 0] = np.clip(add[:, 0], 0, sensor_w - 1)
    add[:, 1] = np.clip(add[:, 1], 0, sensor_h - 1)
    out = np.concatenate((events, add))
    return out[out[:, 2].argsort()]


def events_augment(args, events, size, seed=None):
    """dataset/augmentation/events_augment.py:80-86."""
    if seed is not None:
        np.random.seed(seed)
    return erase_and_add_events(args, events, size=size)
