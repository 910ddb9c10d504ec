from .sim_facade import IsaacGymEnv, TerrainGymEnv
from .vec_env import ShifuVecEnv
