"""Class-attribute configuration trees (interface mirror of ``shifu/configs``).

A config is a class whose attributes are values or nested classes; instantiating it turns every
nested class into an instance, recursively (``shifu/configs/base_config.py:37-58``), so user code
can write ``cfg.terrain.num_rows = 3`` on an instance without touching the class.  The ``sim`` /
``asset_options`` namespaces are copied onto the matching ``gymapi`` structs
(``env_config.py:22-36``, ``asset_config.py:22-30``).

The hot path reads its launch constants from these objects once, at env construction
(gains, limits, grid, scales, thresholds -> ``ShifuA1Desc`` / ``ShifuAbbDesc`` in
``include/shifu_b200.h``).  Default values are the reference's (cited per field).
"""
from __future__ import annotations

import inspect

from isaacgym import gymapi


def _public_fields(ns):
    return [(k, getattr(ns, k)) for k in dir(ns) if "__" not in k]


class BaseConfig:
    name = None

    def __init__(self) -> None:
        self.init_member_classes(self)

    @staticmethod
    def init_member_classes(obj):
        for key in dir(obj):
            if key == "__class__":
                continue
            val = getattr(obj, key)
            if inspect.isclass(val):
                inst = val()
                setattr(obj, key, inst)
                BaseConfig.init_member_classes(inst)


# ---------------------------------------------------------------------------------------------
# environments (shifu/configs/env_config.py)
# ---------------------------------------------------------------------------------------------


class BaseEnvConfig(BaseConfig):
    num_envs = 5
    num_obs = 10
    num_privileged_obs = None
    num_actions = 3
    num_actions_history = None
    send_timeouts = True
    episode_length_s = 20
    spacing = 1.
    device = 'cuda:0'
    physics_engine = gymapi.SIM_PHYSX

    def __init__(self):
        self._init_sim_params()
        super().__init__()

    def _init_sim_params(self):
        params = gymapi.SimParams()
        for key, val in _public_fields(self.sim):
            if key == 'physx':
                for pk, pv in _public_fields(val):
                    setattr(params.physx, pk, pv)
            else:
                setattr(params, key, val)
        self.sim_params = params

    class sim:
        dt = 0.005
        substeps = 1
        up_axis = gymapi.UP_AXIS_Z
        gravity = gymapi.Vec3(0.0, 0.0, -9.81)
        use_gpu_pipeline = True

        class physx:
            num_threads = 10
            use_gpu = True
            solver_type = 1
            num_position_iterations = 8
            num_velocity_iterations = 1
            contact_offset = 0.01
            rest_offset = 0.0
            bounce_threshold_velocity = 0.5
            max_depenetration_velocity = 1.0
            max_gpu_contact_pairs = 2 ** 23
            default_buffer_size_multiplier = 5

    class debug:
        headless = False
        camera_pos = [1., -1., 1.]
        camera_lookat = [0, 0, 0]
        enable_viewer_sync = True
        viewer_attach_robot_env_idx = None

    class normalization:
        clip_observations = 100.
        clip_actions = 1.

    class control:
        decimation = 4


class TerrainEnvConfig(BaseEnvConfig):
    class terrain:                      # env_config.py:77-102
        mesh_type = 'trimesh'
        horizontal_scale = 0.1
        vertical_scale = 0.005
        border_size = 25
        static_friction = 1.0
        dynamic_friction = 1.0
        restitution = 0.
        measure_heights = True
        measured_points_x = [-0.8, -0.7, -0.6, -0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5,
                             0.6, 0.7, 0.8]
        measured_points_y = [-0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5]
        selected = False
        terrain_kwargs = None
        terrain_length = 8.
        terrain_width = 8.
        num_rows = 10
        num_cols = 20
        terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2]
        slope_treshold = 0.75
        curriculum = True
        max_init_terrain_level = 5


# ---------------------------------------------------------------------------------------------
# actors (shifu/configs/asset_config.py)
# ---------------------------------------------------------------------------------------------


class ActorConfig(BaseConfig):
    name = "DummyActor"
    root_dir = "./asset"
    urdf_filename = None
    default_pos = [0, 0, 0]
    default_quat = [0, 0, 0, 1]
    default_dof_pos = None
    domain_randomization = False
    dof_stiffness = None
    dof_damping = None

    def __init__(self):
        self._init_asset_options()
        super().__init__()

    def _init_asset_options(self):
        opts = gymapi.AssetOptions()
        for key, val in _public_fields(self.asset_options):
            setattr(opts, key, val)
        self.asset_options = opts

    class asset_options:
        fix_base_link = False
        default_dof_drive_mode = gymapi.DOF_MODE_NONE
        disable_gravity = False
        collapse_fixed_joints = True
        flip_visual_attachments = False
        replace_cylinder_with_capsule = False
        mesh_normal_mode = gymapi.FROM_ASSET
        use_physx_armature = True
        thickness = 0.001


class BoxActorConfig(ActorConfig):
    box_dim = [0.05, 0.05, 0.05]
    mass = 0.1
    friction = 0.5
    color = [1., 1., 1.]

    class rigid_shape_props:
        friction = 1.0
        torsion_friction = 0.001
        restitution = 0.0


class ArmRobotActorConfig(ActorConfig):
    name = "DummyArmActor"
    default_dof_pos = [0., 0., 0.5]
    end_effector_names = ['tip0']
    dof_stiffness = [400] * 5
    dof_damping = [80] * 5
    end_effector_velocity = 0.1
    min_ee_pos = [-0.25, -0.25, 0.11]
    max_ee_pos = [0.25, 0.25, 0.14]

    class asset_options(ActorConfig.asset_options):
        fix_base_link = True
        default_dof_drive_mode = gymapi.DOF_MODE_POS
        disable_gravity = True
        replace_cylinder_with_capsule = True


class LeggedRobotActorConfig(ActorConfig):
    name = "DummyArmActor"
    default_dof_pos = [0.] * 12
    end_effector_names = ['foot0', 'foot1', 'foot2', 'foot3']
    dof_stiffness = [20] * 12
    dof_damping = [.5] * 12

    class asset_options(ActorConfig.asset_options):
        fix_base_link = False
        default_dof_drive_mode = gymapi.DOF_MODE_POS
        disable_gravity = False
        replace_cylinder_with_capsule = True
        flip_visual_attachments = True


# ---------------------------------------------------------------------------------------------
# sensors / policy: carried for import compatibility only — camera rendering and PPO are
# outside the hot path.
# ---------------------------------------------------------------------------------------------


class BaseSensorConfig(BaseConfig):
    name = "DummySensor"
    frequency = 30
    data_shape = 0


class CameraSensorConfig(BaseSensorConfig):
    name = "DummyCameraSensor"
    image_types = [gymapi.IMAGE_COLOR, gymapi.IMAGE_DEPTH, gymapi.IMAGE_SEGMENTATION,
                   gymapi.IMAGE_OPTICAL_FLOW]
    image_normalization = False
    local_lookat_positions = None
    transform = None
    attach_local_transform = None

    def __init__(self):
        props = gymapi.CameraProperties()
        for key, val in _public_fields(self.camera_props):
            setattr(props, key, val)
        self.camera_props = props
        super().__init__()

    class camera_props:
        enable_tensors = True


class PPOConfig(BaseConfig):
    seed = 1
    runner_class_name = 'OnPolicyRunner'

    class policy:
        init_noise_std = 1.0
        actor_hidden_dims = [512, 256, 128]
        critic_hidden_dims = [512, 256, 128]
        activation = 'elu'

    class algorithm:
        value_loss_coef = 1.0
        use_clipped_value_loss = True
        clip_param = 0.2
        entropy_coef = 0.01
        num_learning_epochs = 5
        num_mini_batches = 4
        learning_rate = 1.e-3
        schedule = 'adaptive'
        gamma = 0.99
        lam = 0.95
        desired_kl = 0.01
        max_grad_norm = 1.

    class runner:
        policy_class_name = 'ActorCritic'
        algorithm_class_name = 'PPO'
        num_steps_per_env = 24
        max_iterations = 1500
        save_interval = 50
        experiment_name = 'test'
        run_name = ''
        resume = False
        load_run = -1
        checkpoint = -1
        resume_path = None
