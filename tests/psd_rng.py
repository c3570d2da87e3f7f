"""numpy restatement of the counter-based input generator (oracle/psdo_common.hpp,
csrc/psd_rng.cuh): splitmix64 keyed by (seed, problem, factor, row, col, part)."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _sm64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def gen_uniform(seed, n, p, batch, first_b=0, part=0):
    """returns [batch][p][col][row] float64 (column-major factors)"""
    with np.errstate(over="ignore"):
        b = (np.arange(batch, dtype=np.uint64) + np.uint64(first_b)).reshape(-1, 1, 1, 1)
        j = np.arange(p, dtype=np.uint64).reshape(1, -1, 1, 1)
        c = np.arange(n, dtype=np.uint64).reshape(1, 1, -1, 1)
        r = np.arange(n, dtype=np.uint64).reshape(1, 1, 1, -1)
        k = _sm64(np.uint64(seed))
        k = _sm64(k ^ (b * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x1234567)))
        k = _sm64(k ^ (j * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(0x89ABCDE)))
        k = _sm64(k ^ ((r << np.uint64(32)) | (c << np.uint64(1)) | np.uint64(part)))
    return (k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
