"""Runner helpers — interface mirror of ``shifu/runner/utils.py`` (log directories, config flattening
for rsl_rl, seeding, checkpoint lookup)."""
from __future__ import annotations

import os
import random
from datetime import datetime

import numpy as np
import torch


def datetime_logdir(log_root: str, run_name: str) -> str:
    """``<log_root>/<Mon01_12-30-00>_<run_name>`` (utils.py:8-11)."""
    return os.path.join(log_root, datetime.now().strftime('%b%d_%H-%M-%S') + '_' + run_name)


def latest_logdir(log_root: str) -> str:
    runs = sorted(d for d in os.listdir(log_root) if os.path.isdir(os.path.join(log_root, d)))
    if not runs:
        raise ValueError(f"No runs in this directory: {log_root}")
    return os.path.join(log_root, runs[-1])


def class_to_dict(obj) -> dict:
    """Config tree -> nested dict for ``OnPolicyRunner`` (utils.py:23-38): public attributes only,
    objects with a ``__dict__`` recursed into, lists element-wise."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        if callable(val) and not isinstance(val, type) and not hasattr(val, "__dict__"):
            continue
        out[key] = [class_to_dict(v) for v in val] if isinstance(val, list) else class_to_dict(val)
    return out


def get_load_path(root: str, load_run=-1, checkpoint=-1) -> str:
    """Latest (or named) run directory and highest-numbered (or named) ``model_*.pt`` in it
    (utils.py:41-60)."""
    run_dir = latest_logdir(root) if load_run == -1 else os.path.join(root, load_run)
    if checkpoint == -1:
        models = sorted((f for f in os.listdir(run_dir) if 'model' in f), key=lambda m: '{0:0>15}'.format(m))
        if not models:
            raise ValueError(f"No model file in {run_dir}")
        model = models[-1]
    else:
        model = "model_{}.pt".format(checkpoint)
    return os.path.join(run_dir, model)


def set_seed(seed: int):
    """python / numpy / torch (+cuda) generators; ``-1`` draws a random seed (utils.py:63-73)."""
    if seed == -1:
        seed = np.random.randint(0, 10000)
    print("Setting seed: {}".format(seed))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    torch.cuda.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    return seed
