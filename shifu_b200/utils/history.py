"""HistoryRecorder — interface mirror of ``shifu/utils/train.py``: an (N, A, H) shift buffer with
the newest sample in slot 0.  ``add`` is the ``shifu_history_add`` kernel; the fused A1 step does
the push inside K-main and only shares ``history_buf`` with this object."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class HistoryRecorder:
    def __init__(self, shape: Sequence[int], num_history: int, device, kernels: Optional[Callable] = None):
        if not (isinstance(num_history, int) and num_history > 0):
            raise AssertionError("num_history must be a positive int")
        self.dshape, self.num_history, self.device = shape, num_history, device
        self.history_buf = torch.zeros((*shape, num_history), device=device)
        self._kernels = kernels          # callable returning an EnvKernels

    def add(self, x: torch.Tensor):
        """Shift every slot one step into the past and store ``x`` as the newest (train.py:12-14)."""
        if self._kernels is None:
            raise RuntimeError("HistoryRecorder.add needs the CUDA kernels (no torch fallback)")
        self._kernels().history_add(self.history_buf, x)

    def reset_idx(self, idx: torch.Tensor):
        self.history_buf[idx] = 0.

    def get_last(self, idx: int):
        """The sample ``idx`` steps ago (0 = newest)."""
        return self.history_buf.select(-1, idx)

    def flatten(self):
        """(N, A, H) -> (N, H*A), slot-major: [newest A values, previous A values, ...] (train.py:33-35)."""
        buf = self.history_buf
        slot_major = buf.permute(0, *range(buf.dim() - 1, 0, -1))
        lead, width = self.dshape[:-1], self.dshape[-1]
        return slot_major.reshape(*lead, width * self.num_history)
