from di_hpc_b200.rl_utils.ppo import *  # noqa: F401,F403
from di_hpc_b200.rl_utils.ppo import PPO, PPOFunction, hpc_ppo_info, hpc_ppo_loss  # noqa: F401
