from .ppo import PPO
from .trpo import TRPO
from .vpg import VPG

__all__ = ["PPO", "TRPO", "VPG"]
