"""``run_policy`` / ``load_policy`` / ``build_policy_runner`` — interface mirror of
``shifu/runner/policy_runner.py`` (the caller of the hot path: rsl_rl's ``OnPolicyRunner`` steps the
env 24 times per iteration, ``shifu/configs/policy_config.py:36``).

PPO itself lives in the external ``rsl_rl`` package, which is imported lazily: the ``random`` mode
(the reference's manual integration check, policy_runner.py:33-41) needs no learner."""
from __future__ import annotations

import torch

from .utils import class_to_dict, datetime_logdir, get_load_path, set_seed


def _on_policy_runner_class():
    from rsl_rl.runners import OnPolicyRunner

    class _OnPolicyRunner(OnPolicyRunner):
        def load(self, path, load_optimizer=True):                      # policy_runner.py:7-14
            state = torch.load(path, map_location=self.device)          # checkpoints move between devices
            self.alg.actor_critic.load_state_dict(state['model_state_dict'])
            if load_optimizer:
                self.alg.optimizer.load_state_dict(state['optimizer_state_dict'])
            self.current_learning_iteration = state['iter']
            return state['infos']

    return _OnPolicyRunner


def build_policy_runner(env, train_cfg, log_root="./logs", device="cuda:0", resume=False):
    log_dir = datetime_logdir(log_root, train_cfg.runner.run_name)
    set_seed(train_cfg.seed)
    runner = _on_policy_runner_class()(env, class_to_dict(train_cfg), log_dir, device=device)
    if resume:
        path = get_load_path(log_root, load_run=train_cfg.runner.load_run, checkpoint=train_cfg.runner.checkpoint)
        print(f"Loading model from: {path}")
        runner.load(path)
    return runner


def load_policy(env, policy_cfg, log_root, device='cuda:0'):
    return build_policy_runner(env, policy_cfg, log_root, resume=True, device=device).get_inference_policy()


def run_policy(run_mode, env_class, env_cfg, policy_cfg, log_root="./logs", play_num_envs=50, play_iterations=3000):
    """``train`` | ``play`` | ``random`` (policy_runner.py:17-43)."""
    if run_mode == 'train':
        env = env_class(env_cfg)
        runner = build_policy_runner(env, policy_cfg, log_root)
        runner.learn(num_learning_iterations=policy_cfg.runner.max_iterations, init_at_random_ep_len=True)
        return env
    if run_mode not in ('play', 'random'):
        raise NotImplementedError(run_mode)
    env_cfg.num_envs = play_num_envs
    env_cfg.debug.headless = False
    env = env_class(env_cfg)
    if run_mode == 'play':
        policy = load_policy(env, policy_cfg, log_root)
        env.reset()
        obs = env.get_observations()
        for _ in range(play_iterations):
            obs, _, _, _, _ = env.step(policy(obs.detach()).detach())
    else:
        env.reset()
        for _ in range(play_iterations):
            actions = 2 * torch.rand(env.num_envs, env.num_actions, device=env.device) - 1
            env.step(actions.detach())
    return env
