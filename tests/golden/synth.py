"""Exact, platform-independent synthetic float inputs shared by make_golden.py and the tests.

Large float inputs are not stored in the fixtures: both sides regenerate them from an integer
hash whose top 24 bits map to an exactly-representable fp32 value in [-0.5, 0.5).
"""
import numpy as np


def hash_uniform(shape, seed):
    n = int(np.prod(shape))
    i = np.arange(n, dtype=np.uint64)
    mixed = np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    h = i * np.uint64(2654435761) + mixed          # uint64 arithmetic wraps mod 2**64
    h ^= h >> np.uint64(29)
    h = h * np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    top = (h >> np.uint64(40)).astype(np.int64)          # 24 bits
    return ((top.astype(np.float64) / float(1 << 24)) - 0.5).astype(np.float32).reshape(shape)
