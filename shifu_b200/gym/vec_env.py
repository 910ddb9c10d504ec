"""ShifuVecEnv — the rsl_rl-compatible vectorised env (interface mirror of ``shifu/gym/env.py``).

Upward contract (``rsl_rl.env.VecEnv``): ``num_envs, num_obs, num_privileged_obs, num_actions,
max_episode_length, episode_length_buf, obs_buf, rew_buf, reset_buf, extras, device``;
``step(actions) -> (obs, privileged_obs, rew, dones, extras)``, ``reset()``,
``get_observations()``, ``get_privileged_observations()``.

User hooks (README.md:95-128): ``build_reward_functions``, ``compute_observations``,
``compute_termination``, optional ``reset_idx`` / ``episode_log`` / ``step``.

Two execution modes, both CUDA-only (no CPU fallback):

* **user-hook mode** (default): the hooks are arbitrary torch code on CUDA tensors and this class
  runs the reference's orchestration (env.py:85-130) around them, with the rows it owns itself —
  action / obs clip, reset-id compaction, history push — done by the ``libshifu_b200`` kernels;
* **fused mode**: a task that declares its reward terms to the registry
  (``shifu_b200.hotpath.compile_reward_terms``) replaces ``post_step`` by one fused kernel
  (see ``shifu_b200/tasks``).
"""
from __future__ import annotations

import typing

import numpy as np
import torch
from rsl_rl.env import VecEnv

from shifu_b200.configs import BaseEnvConfig, TerrainEnvConfig
from shifu_b200.gym.sim_facade import IsaacGymEnv, TerrainGymEnv
from shifu_b200.utils.history import HistoryRecorder


class ShifuVecEnv(VecEnv):
    def __init__(self, cfg: BaseEnvConfig, env_offset: int = 0, num_envs_global: int = None):
        self.cfg = cfg
        if isinstance(cfg, TerrainEnvConfig):
            self.isg_env = TerrainGymEnv(cfg, env_offset=env_offset, num_envs_global=num_envs_global)
        else:
            self.isg_env = IsaacGymEnv(cfg)
        self.env_offset = int(env_offset)
        self.num_envs_global = int(num_envs_global) if num_envs_global else int(cfg.num_envs)
        self.num_envs = self.isg_env.num_envs
        self.device = self.isg_env.device
        self.num_obs = cfg.num_obs
        self.num_privileged_obs = cfg.num_privileged_obs
        self.num_actions = cfg.num_actions
        self.clip_obs = cfg.normalization.clip_observations
        self.clip_actions = cfg.normalization.clip_actions
        self.max_episode_length_s = cfg.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.isg_env.dt)

        n, dev = self.num_envs, self.device
        self.actions = torch.zeros(n, self.num_actions, device=dev, dtype=torch.float, requires_grad=False)
        self.obs_buf = torch.zeros(n, self.num_obs, device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(n, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(n, device=dev, dtype=torch.long)
        self._episode_length_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.time_out_buf = torch.zeros(n, device=dev, dtype=torch.bool)
        self.extras = {}
        if self.cfg.num_actions_history:
            self.actions_recorder = HistoryRecorder(self.actions.shape, self.cfg.num_actions_history, device=dev,
                                                    kernels=self.isg_env.kernels)
        self.privileged_obs_buf = None if self.num_privileged_obs is None \
            else torch.zeros(n, self.num_privileged_obs, device=dev, dtype=torch.float)
        self.common_step_counter = 0
        self.stats_allreduce = None            # callable(tensor) summing over ranks (sharded runs)
        self.reward_functions = self.build_reward_functions()
        self._prepare_reward_functions()

    # rsl_rl re-binds ``env.episode_length_buf = randint_like(...)`` (init_at_random_ep_len,
    # policy_runner.py:22); the kernels hold the buffer's address, so assignment copies in place.
    @property
    def episode_length_buf(self):
        return self._episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        if value is self._episode_length_buf:
            return
        self._episode_length_buf.copy_(torch.as_tensor(value, device=self.device))

    def destroy(self):
        self.isg_env.destroy()

    # -- user hooks ----------------------------------------------------------------------------
    def build_reward_functions(self) -> typing.List:
        raise NotImplementedError

    def compute_observations(self):
        raise NotImplementedError

    def compute_termination(self):
        raise NotImplementedError

    def episode_log(self, env_ids) -> typing.Dict:
        pass

    # -- orchestration (user-hook mode) ----------------------------------------------------------
    def step(self, actions: torch.Tensor):
        assert self.isg_env.robot, "add robot before step"
        k = self.isg_env.kernels()
        self.actions = k.clip(actions, self.clip_actions, out=self.actions)      # env.py:87
        self.isg_env.step(self.actions)
        self.post_step()
        self.obs_buf = k.clip(self.obs_buf, self.clip_obs)                       # env.py:90
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def post_step(self):
        self._episode_length_buf += 1
        self.common_step_counter += 1
        self.compute_termination()
        self.compute_reward()
        env_ids = self.isg_env.kernels().nonzero(self.reset_buf)                 # env.py:101
        self.reset_idx(env_ids)
        self.compute_observations()
        self.isg_env.refresh_sensors()
        if self.cfg.num_actions_history:
            self.actions_recorder.add(self.actions)

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, pri_obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device,
                                                      requires_grad=False))
        return obs, pri_obs

    def reset_idx(self, env_ids):
        if len(env_ids) == 0:
            return
        self.isg_env.reset_idx(env_ids)
        self._episode_length_buf[env_ids] = 0
        self.reset_buf[env_ids] = 1
        if self.cfg.num_actions_history:
            self.actions_recorder.reset_idx(env_ids)
        self.extras["episode"] = {}
        self.log_info(env_ids)
        if self.cfg.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    def log_info(self, env_ids):
        for key in self.episode_rewards.keys():
            self.extras["episode"][key] = torch.mean(self.episode_rewards[key][env_ids]) / self.max_episode_length_s
            self.episode_rewards[key][env_ids] = 0.
        ep_info = self.episode_log(env_ids)
        if ep_info:
            self.extras["episode"].update(ep_info)

    def _prepare_reward_functions(self):
        assert len(self.reward_functions) > 0
        self.episode_rewards = {fn.__name__: torch.zeros(self.num_envs, device=self.device, dtype=torch.float)
                                for fn in self.reward_functions}

    def compute_reward(self):
        self.rew_buf[:] = 0.
        for rew_func in self.reward_functions:
            rew = rew_func()
            self.episode_rewards[rew_func.__name__] += rew
            self.rew_buf[:] += rew

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf
