"""ctypes binding of oracle/ep_oracle.c (test infrastructure; see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libep_oracle.so")
_lib = None

EINVAL, EINDEX = -1, -2


def build(force=False):
    src = os.path.join(_HERE, "ep_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libep_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _raise(rc, what):
    if rc == EINDEX:
        raise IndexError(f"{what}: flat index out of range (the reference raises here)")
    if rc != 0:
        raise ValueError(f"{what}: invalid arguments (rc={rc})")


def _aos(events):
    ev = np.ascontiguousarray(events)
    assert ev.ndim == 2 and ev.shape[1] == 4
    if ev.dtype not in (np.float64, np.float32):
        ev = ev.astype(np.float64)
    return ev


def voxel_grid(events, num_bins, size):
    ev = _aos(events)
    H, W = int(size[0]), int(size[1])
    out = np.empty((num_bins, H, W), np.float32)
    fn = lib().oracle_voxel_grid_f64 if ev.dtype == np.float64 else lib().oracle_voxel_grid_f32
    rc = fn(_ptr(ev), ctypes.c_int64(ev.shape[0]), num_bins, H, W, _ptr(out))
    _raise(rc, "voxel_grid")
    return out


def count_frame(events, size, channels=2):
    ev = _aos(events)
    H, W = int(size[0]), int(size[1])
    out = np.empty((channels, H, W), np.float32)
    fn = lib().oracle_count_frame_f64 if ev.dtype == np.float64 else lib().oracle_count_frame_f32
    rc = fn(_ptr(ev), ctypes.c_int64(ev.shape[0]), H, W, channels, _ptr(out))
    _raise(rc, "count_frame")
    return out


def remove_hot_pixel_mem(hist, num_stds=10):
    h = np.ascontiguousarray(hist, np.float32).copy()
    assert h.ndim == 3 and h.shape[0] == 3
    lib().oracle_remove_hot_pixel_mem(_ptr(h), h.shape[1], h.shape[2], ctypes.c_float(num_stds))
    return h


def count_normalise(img):
    h = np.ascontiguousarray(img, np.float32).copy()
    lib().oracle_count_normalise(_ptr(h), h.shape[0], h.shape[1], h.shape[2])
    return h


def mem_normalise(img, guard_zero=False):
    h = np.ascontiguousarray(img, np.float32).copy()
    lib().oracle_mem_normalise(_ptr(h), h.shape[1], h.shape[2], int(guard_zero))
    return h


def evrep(xs, ys, ts, ps, resolution):
    W, H = int(resolution[0]), int(resolution[1])
    xs = np.ascontiguousarray(xs, np.int16)
    ys = np.ascontiguousarray(ys, np.int16)
    ts = np.ascontiguousarray(ts, np.float64)
    ps = np.ascontiguousarray(ps, np.float64)
    out = np.empty((3, H, W), np.float64)
    rc = lib().oracle_evrep(_ptr(xs), _ptr(ys), _ptr(ts), _ptr(ps), ctypes.c_int64(xs.shape[0]), W, H, _ptr(out))
    _raise(rc, "evrep")
    return out


def voxel_grid_batch(events, offsets, num_bins, size, num_threads=1):
    ev = np.ascontiguousarray(events, np.float64)
    off = np.ascontiguousarray(offsets, np.int64)
    B = off.shape[0] - 1
    H, W = int(size[0]), int(size[1])
    out = np.empty((B, num_bins, H, W), np.float32)
    rc = lib().oracle_voxel_grid_batch_f64(_ptr(ev), _ptr(off), B, num_bins, H, W, _ptr(out), int(num_threads))
    _raise(rc, "voxel_grid_batch")
    return out


def count_frame_batch(events, offsets, size, channels=2, num_threads=1):
    ev = np.ascontiguousarray(events, np.float64)
    off = np.ascontiguousarray(offsets, np.int64)
    B = off.shape[0] - 1
    H, W = int(size[0]), int(size[1])
    out = np.empty((B, channels, H, W), np.float32)
    rc = lib().oracle_count_frame_batch_f64(_ptr(ev), _ptr(off), B, H, W, channels, _ptr(out), int(num_threads))
    _raise(rc, "count_frame_batch")
    return out
