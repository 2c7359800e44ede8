from typing import List, Optional, Tuple

import numpy as np


class Evaluator:
    """Runs evaluation episodes on a gymnasium-protocol env (ref: evaluator.py:9-52); host-side, env-bound."""

    def __init__(self, seed: Optional[int] = None):
        self.seed = seed

    def evaluate(self, policy, env, num_episodes: int) -> Tuple[List[float], List[int]]:
        returns: List[float] = []
        lengths: List[int] = []
        observation, _ = env.reset(seed=self.seed)
        for _ in range(num_episodes):
            total, steps, finished = 0.0, 0, False
            while not finished:
                action: np.ndarray = policy.get_action_numpy(observation)
                observation, reward, terminated, truncated, _ = env.step(action)
                finished = terminated or truncated
                total += reward
                steps += 1
            observation, _ = env.reset()
            returns.append(total)
            lengths.append(steps)
        return returns, lengths
