"""Drop-in namespace: ``import shifu`` resolves to the B200-native implementation (``shifu_b200``), so
the reference's examples (``from shifu.units import LeggedRobot``, ``from shifu.gym import ShifuVecEnv``,
``from shifu.runner import run_policy``, ``from shifu.configs import ...``) run unchanged.

The simulator package (``isaacgym``) must be importable first — the real one on a simulator box, or
the stand-in (``shifu_b200.sim.fake_isaacgym.install()``) where only the hot path is exercised.
"""
import importlib
import sys

_ALIASES = {
    # reference module -> implementation module
    "shifu.configs": "shifu_b200.configs", "shifu.units": "shifu_b200.units", "shifu.gym": "shifu_b200.gym",
    "shifu.runner": "shifu_b200.runner", "shifu.utils": "shifu_b200.utils",
    "shifu.gym.env": "shifu_b200.gym.vec_env", "shifu.gym.isaac_gym": "shifu_b200.gym.sim_facade",
    "shifu.units.units": "shifu_b200.units.base", "shifu.units.robot": "shifu_b200.units.robots",
    "shifu.units.object": "shifu_b200.units.objects", "shifu.units.sensors": "shifu_b200.units.sensors",
    "shifu.utils.train": "shifu_b200.utils.history", "shifu.utils.terrain": "shifu_b200.utils.heightmap",
    "shifu.runner.utils": "shifu_b200.runner.utils", "shifu.runner.policy_runner": "shifu_b200.runner.policy_runner",
    "shifu.configs.base_config": "shifu_b200.configs.config_tree", "shifu.configs.env_config": "shifu_b200.configs.config_tree",
    "shifu.configs.asset_config": "shifu_b200.configs.config_tree", "shifu.configs.sensor_config": "shifu_b200.configs.config_tree",
    "shifu.configs.policy_config": "shifu_b200.configs.config_tree",
}

for _ref, _impl in _ALIASES.items():
    _mod = importlib.import_module(_impl)
    sys.modules[_ref] = _mod
    _parent, _, _leaf = _ref.rpartition(".")
    if _parent == "shifu":
        globals()[_leaf] = _mod
