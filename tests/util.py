"""Shared helpers for the parity tests."""
import numpy as np

P = 0xFFFFFFFF00000001


def splitmix64(seed, n):
    """SplitMix64 stream (SURVEY.md §8(d) synthetic inputs), vectorised."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand_felts(seed, shape, canonical=True):
    n = int(np.prod(shape))
    v = splitmix64(seed, n)
    if canonical:
        v = np.where(v >= np.uint64(P), v - np.uint64(P), v)
    return v.reshape(shape)


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r
