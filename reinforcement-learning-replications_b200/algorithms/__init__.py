from .ppo import PPO
from .td3 import DDPG, TD3
from .trpo import TRPO
from .vpg import VPG

__all__ = ["VPG", "TRPO", "PPO", "DDPG", "TD3"]
