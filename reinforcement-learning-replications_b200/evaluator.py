"""Evaluation roll-outs on a gymnasium-protocol environment (ref: evaluator.py:9-52).  Host-side and env-bound: outside
the GPU path, kept because ``TD3.learn`` / ``DDPG.learn`` report evaluation returns every epoch."""
from itertools import count
from typing import List, Optional, Tuple


class Evaluator:
    def __init__(self, seed: Optional[int] = None):
        self.seed = seed

    @staticmethod
    def _play(policy, env, first_observation):
        """One episode from ``first_observation``: (sum of rewards, number of steps)."""
        observation, episode_return = first_observation, 0.0
        for step in count(1):
            observation, reward, terminated, truncated, _ = env.step(policy.get_action_numpy(observation))
            episode_return += reward
            if terminated or truncated:
                return episode_return, step

    def evaluate(self, policy, env, num_episodes: int) -> Tuple[List[float], List[int]]:
        """(episode returns, episode lengths) of ``num_episodes`` consecutive episodes; only the first reset is seeded,
        like the reference's."""
        results = []
        start, _ = env.reset(seed=self.seed)
        for _ in range(num_episodes):
            results.append(self._play(policy, env, start))
            start, _ = env.reset()
        return [r for r, _ in results], [n for _, n in results]
