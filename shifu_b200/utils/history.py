"""HistoryRecorder — interface mirror of ``shifu/utils/train.py`` ((N, A, H) shift buffer, newest
sample in slot 0).  ``add`` is the ``shifu_history_add`` kernel; the fused A1 step does the push
inside K-main and only shares ``history_buf`` with this object."""
from __future__ import annotations

import torch


class HistoryRecorder:
    def __init__(self, shape, num_history, device, kernels=None):
        assert isinstance(num_history, int) and num_history > 0, "num_history must be a positive int"
        self.dshape = shape
        self.num_history = num_history
        self.device = device
        self.history_buf = torch.zeros(*shape, num_history, device=device)
        self._kernels = kernels          # callable returning an EnvKernels

    def add(self, x):
        if self._kernels is None:
            raise RuntimeError("HistoryRecorder.add needs the CUDA kernels (no torch fallback)")
        self._kernels().history_add(self.history_buf, x)

    def reset_idx(self, idx):
        self.history_buf.index_fill_(0, idx, 0.)

    def get_last(self, idx):
        """t-idx sample (0 = newest)."""
        return self.history_buf[..., idx]

    def flatten(self):
        nd = len(self.history_buf.shape)
        p = self.history_buf.permute(0, *reversed(range(1, nd)))
        return p.reshape(*self.dshape[:-1], self.dshape[-1] * self.num_history)
