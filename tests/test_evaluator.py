"""Evaluator (host-side): episode returns / lengths and the reset protocol of the reference (evaluator.py:25-52): the
first reset is seeded, every later one is not, and each episode runs until terminated OR truncated."""
import numpy as np

from rl_replicas_b200.evaluator import Evaluator


class _Env:
    """Episode k lasts k + 2 steps (odd episodes end by truncation); reward = step index."""

    def __init__(self):
        self.resets, self.episode, self.t = [], -1, 0

    def reset(self, seed=None):
        self.resets.append(seed)
        self.episode += 1
        self.t = 0
        return np.asarray([float(self.episode)]), {}

    def step(self, action):
        assert action.shape == (1,)
        self.t += 1
        last = self.t == self.episode + 2
        return np.asarray([float(self.t)]), float(self.t), last and self.episode % 2 == 0, last and self.episode % 2 == 1, {}


class _Policy:
    def get_action_numpy(self, observation):
        return np.asarray(observation)


def test_returns_lengths_and_reset_protocol():
    env = _Env()
    returns, lengths = Evaluator(seed=7).evaluate(_Policy(), env, 4)
    assert lengths == [2, 3, 4, 5]
    assert returns == [3.0, 6.0, 10.0, 15.0]
    assert env.resets == [7, None, None, None, None]  # one reset per episode after the seeded first one
    assert all(isinstance(r, float) for r in returns) and all(isinstance(n, int) for n in lengths)
