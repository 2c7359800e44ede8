from .ppo import PPO
from .vpg import VPG

__all__ = ["PPO", "VPG"]
