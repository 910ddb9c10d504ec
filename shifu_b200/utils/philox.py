"""Philox4x32-10 on torch int64 tensors (any device) for USER-HOOK tasks that want the same
counter-based reset draws as the fused kernels (``csrc/philox.cuh``): counter =
(env_id_global, step, stream, 0), key = (seed_lo, seed_hi).  The fused path never calls this."""
from __future__ import annotations

import torch

_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def _mulhilo(m: int, c: torch.Tensor):
    """(hi32, lo32) of m*c for a 32-bit constant m and c in [0, 2^32) held in int64."""
    p_lo = m * (c & 0xFFFF)
    p_hi = m * (c >> 16)
    low = p_lo + ((p_hi & 0xFFFF) << 16)
    return (p_hi >> 16) + (low >> 32), low & _MASK


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    for _ in range(10):
        hi0, lo0 = _mulhilo(_M0, c0)
        hi1, lo1 = _mulhilo(_M1, c2)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def _lanes(seed: int, env_ids: torch.Tensor, step: int, stream: int):
    e = env_ids.to(torch.int64) & _MASK
    z = torch.zeros_like(e)
    return philox4x32_10(e, z + (step & _MASK), z + stream, z, seed & _MASK, (seed >> 32) & _MASK)


def draw_u01(seed: int, env_ids: torch.Tensor, step: int, stream: int, lanes: int) -> torch.Tensor:
    """(R, lanes) float32 in [0,1) on the 24-bit grid: (x >> 8) * 2^-24."""
    out = _lanes(seed, env_ids, step, stream)[:lanes]
    return torch.stack([(x >> 8).to(torch.float32) * (2.0 ** -24) for x in out], dim=1)


def draw_randint(seed: int, env_ids: torch.Tensor, step: int, stream: int, high: int) -> torch.Tensor:
    x = _lanes(seed, env_ids, step, stream)[0]
    return ((x >> 8) * high) >> 24
