"""ORACLE (test infrastructure, not product code): Philox4x32-10 in numpy.

Counter-based generator of Salmon et al., "Parallel random numbers: as easy as
1, 2, 3" (SC'11).  Pinned against the Random123 known-answer vectors in
``tests/test_philox.py``.  The reference has no counter-based RNG: its reset-time
draws come from torch's / numpy's global generators consumed in a data-dependent
order.  Parity is therefore defined on *injected* draws:
both the reference harness and the CUDA path obtain draw ``(env, step, stream)``
from this function.

Stream map:
  0  terrain-level re-draw for robots that solved the last level (lane 0)
     ``examples/a1_conditional/a1_conditional.py:219``
  1  reset xy offset (lanes 0,1)        ``a1_conditional.py:47``
  2  random push force (lanes 0,1,2)     ``a1_conditional.py:86``
  3  command vx, vy, yaw-rate (lanes 0,1,2)  ``a1_conditional.py:195-199``
  4  ABB cube position (lanes 0,1,2)     ``examples/abb_pushbox_vision/a_prior_stage.py:41``
  5  ABB cube euler (lanes 0,1,2)        ``a_prior_stage.py:44``
  6  ABB goal position                   (same code through ``GoalBox``)
  7  ABB goal euler
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

STREAM_LEVEL, STREAM_XY, STREAM_FORCE, STREAM_CMD = 0, 1, 2, 3
STREAM_CUBE_POS, STREAM_CUBE_EUL, STREAM_GOAL_POS, STREAM_GOAL_EUL = 4, 5, 6, 7


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32 counters; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> _S32).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> _S32).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def draw_u32(seed: int, env_ids, step: int, stream: int):
    """uint32 lanes (4, R) for counter (env_id, step, stream, 0), key = (seed_lo, seed_hi)."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    z = np.zeros_like(env_ids)
    out = philox4x32_10(env_ids, z + np.uint32(step & 0xFFFFFFFF), z + np.uint32(stream), z,
                        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(out, axis=0)


def u01_f32(x_u32):
    """uint32 -> float32 in [0,1) on torch.rand's 24-bit grid: (x >> 8) * 2**-24."""
    return ((x_u32 >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


def u01_f64(x_u32):
    """Same 24-bit grid as a float64 (for the numpy ``np.random.uniform`` call sites)."""
    return ((x_u32 >> np.uint32(8)).astype(np.float64)) * (2.0 ** -24)


def randint10(x_u32, high: int):
    """floor(u * high) for u on the 24-bit grid, done in integers: ((x>>8)*high)>>24."""
    return (((x_u32 >> np.uint32(8)).astype(np.uint64) * np.uint64(high)) >> np.uint64(24)).astype(np.int64)
