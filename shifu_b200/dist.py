"""Multi-GPU plumbing: one process per GPU, envs sharded in contiguous ranges, no data-path
collective — the only exchange is the per-step statistics vector.

* rank r of W owns global envs ``[r*N/W, (r+1)*N/W)``; every rank holds a full replica of the
  height map and the spawn-origin table;
* terrain types (``shifu/gym/isaac_gym.py:342-344``) and the Philox counters use the GLOBAL env
  id, so results are invariant to the GPU count;
* per step each rank contributes ``double[16]`` = [sum over its resetting envs of each term's
  episode sum, #resets, sum of terrain levels over all its envs, sum of successes, #envs] and one
  ``all_reduce(SUM)`` (NCCL over NVLink on GPUs, gloo in the CPU tests) makes the logged means
  global: ``extras["episode"][k] = sum_k / n_reset / max_episode_length_s`` (env.py:149-153).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

from . import _native as nv


def shard_range(num_envs_global: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(env_offset, num_envs_local) of ``rank``; the global count must divide evenly so that every
    GPU runs the same grid (weak scaling: fixed work per GPU)."""
    if num_envs_global % world_size != 0:
        raise ValueError(f"num_envs_global={num_envs_global} must be a multiple of world_size={world_size}")
    n = num_envs_global // world_size
    return rank * n, n


def global_terrain_types(env_offset: int, num_envs_local: int, num_envs_global: int, num_cols: int) -> torch.Tensor:
    """isaac_gym.py:342-344 evaluated on global env ids."""
    gid = torch.arange(env_offset, env_offset + num_envs_local)
    return torch.div(gid, (num_envs_global / num_cols), rounding_mode='floor').to(torch.long)


def make_stats_allreduce(group=None):
    """Callable for ``ShifuVecEnv.stats_allreduce`` / ``A1HotPath.finalize(allreduce=...)``."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None

    def _allreduce(stats: torch.Tensor):
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)

    return _allreduce


def pack_stats(term_sums: Sequence[float], n_reset: int, level_sum: float, success_sum: float,
               n_envs: int) -> torch.Tensor:
    """Host-side constructor of the statistics vector (layout of SHIFU_STAT_* in the header)."""
    s = torch.zeros(nv.NUM_STATS, dtype=torch.double)
    for i, v in enumerate(term_sums):
        s[nv.STAT_TERM0 + i] = v
    s[nv.STAT_NRESET], s[nv.STAT_LEVEL_SUM] = n_reset, level_sum
    s[nv.STAT_SUCCESS], s[nv.STAT_NENVS] = success_sum, n_envs
    return s


def extras_from_stats(stats: torch.Tensor, terms: Sequence[str], max_episode_length_s: float,
                      previous: Dict[str, float] = None) -> Dict[str, float]:
    """What ``shifu_publish_extras`` computes on the device, restated on the host (used by tests and
    by callers that reduce statistics over several steps): values persist when nobody reset."""
    out = dict(previous or {})
    n_reset = float(stats[nv.STAT_NRESET])
    if n_reset <= 0:
        return out
    for i, k in enumerate(terms):
        out[k] = float(stats[nv.STAT_TERM0 + i]) / n_reset / max_episode_length_s
    out["terrain_levels"] = float(stats[nv.STAT_LEVEL_SUM]) / float(stats[nv.STAT_NENVS])
    out["success_rate"] = float(stats[nv.STAT_SUCCESS]) / n_reset
    return out
