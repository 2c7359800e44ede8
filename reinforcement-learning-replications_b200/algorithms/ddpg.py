from .td3 import DDPG

__all__ = ["DDPG"]
