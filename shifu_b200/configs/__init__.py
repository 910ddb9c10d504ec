from .config_tree import (BaseConfig, BaseEnvConfig, TerrainEnvConfig, ActorConfig, BoxActorConfig,
                          ArmRobotActorConfig, LeggedRobotActorConfig, BaseSensorConfig,
                          CameraSensorConfig, PPOConfig)
