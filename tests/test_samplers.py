"""BatchSampler against a hand-rolled rollout -- the reference's own sampler test (ref: tests/test_samplers.py:27-103:
same seed -> same observations / actions / rewards / last observations / dones), on a gymnasium-protocol toy
environment because gymnasium is not installed here.  Also the two behaviours the reference's sampler has beyond that
test: the epoch cut-off closes an episode without setting ``done`` (ref: samplers/batch_sampler.py:65-88) and
``is_continuous`` keeps the running observation across calls (ref: :49-53)."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from rl_replicas_b200.policies import RandomPolicy
from rl_replicas_b200.samplers import BatchSampler, Sampler


class _Space:
    """action_space with gymnasium's ``seed`` / ``sample`` surface."""

    def __init__(self, n):
        self.n, self.rng = n, np.random.default_rng()

    def seed(self, seed):
        self.rng = np.random.default_rng(seed)

    def sample(self):
        return self.rng.uniform(-1.0, 1.0, self.n).astype(np.float32)


class _PoleLikeEnv:
    """Seeded dynamics; an episode terminates when |x| leaves a band or is truncated after 25 steps."""

    def __init__(self):
        self.action_space = _Space(2)
        self.rng = np.random.default_rng()
        self.x, self.t = None, 0
        self.step_noise = 0.3

    def reset(self, seed=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.x = self.rng.uniform(-0.05, 0.05, 3).astype(np.float32)
        self.t = 0
        return self.x.copy(), {}

    def step(self, action):
        self.t += 1
        self.x = (self.x + self.step_noise * self.rng.standard_normal(3) + 0.01 * float(np.sum(action))).astype(np.float32)
        terminated = bool(np.abs(self.x).max() > 1.5)
        truncated = self.t >= 25
        return self.x.copy(), 1.0 + float(self.x[0]), terminated, truncated, {}


def _manual_rollout(seed, num_samples):
    env = _PoleLikeEnv()
    env.action_space.seed(seed)
    out = {k: [] for k in ("observations", "actions", "rewards", "last_observations", "dones")}
    observation, _ = env.reset(seed=seed)
    for step in range(num_samples):
        out["observations"].append(observation)
        action = env.action_space.sample()
        out["actions"].append(action)
        observation, reward, terminated, truncated, _ = env.step(action)
        finished = terminated or truncated
        out["rewards"].append(reward)
        out["dones"].append(finished)
        if finished or step == num_samples - 1:
            out["last_observations"].append(observation)
            if finished:
                observation, _ = env.reset()
    return out


@pytest.mark.parametrize("packed", [False, True])
def test_sample_reproduces_a_manual_rollout(packed):
    seed, n = 0, 1000
    expected = _manual_rollout(seed, n)
    env = _PoleLikeEnv()
    env.action_space.seed(seed)
    sampler: Sampler = BatchSampler(env, seed, packed=packed)
    experience = sampler.sample(n, RandomPolicy(env.action_space))
    assert_array_equal(experience.flattened_observations, expected["observations"])
    assert_array_equal(experience.flattened_actions, expected["actions"])
    assert_array_equal(experience.flattened_rewards, expected["rewards"])
    assert_array_equal(experience.last_observations, expected["last_observations"])
    assert_array_equal(experience.flattened_dones, expected["dones"])
    # bookkeeping the algorithms log from (ref: batch_sampler.py:75-88)
    assert sum(experience.episode_lengths) == n
    assert len(experience.episode_lengths) == len(expected["last_observations"])
    np.testing.assert_allclose(sum(experience.episode_returns), sum(expected["rewards"]), rtol=1e-6)


def test_epoch_cut_off_closes_the_episode_without_done():
    env = _PoleLikeEnv()
    env.action_space.seed(3)
    env.step_noise = 0.0  # nothing can terminate; truncation comes at step 25
    experience = BatchSampler(env, 3).sample(30, RandomPolicy(env.action_space))
    assert list(experience.episode_lengths) == [25, 5]
    assert [bool(d) for d in experience.episode_dones] == [True, False]  # truncated counts as done; the cut-off does not
    flat = [bool(d) for d in experience.flattened_dones]
    assert flat == [False] * 24 + [True] + [False] * 5
    assert len(experience.last_observations) == 2


def test_is_continuous_keeps_the_running_observation():
    def run(is_continuous):
        env = _PoleLikeEnv()
        env.action_space.seed(1)
        sampler = BatchSampler(env, 1, is_continuous=is_continuous)
        policy = RandomPolicy(env.action_space)
        first = sampler.sample(10, policy)
        carried = np.array(sampler.observation, copy=True)
        second = sampler.sample(10, policy)
        return first, carried, second

    _, carried, second = run(True)
    assert_array_equal(np.asarray(second.flattened_observations)[0], carried)  # no reset between calls
    _, carried, second = run(False)
    assert not np.array_equal(np.asarray(second.flattened_observations)[0], carried)  # reset() drew a fresh start
