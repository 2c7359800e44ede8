"""The reference's sampler test (ref: tests/test_samplers.py:27-109) on CartPole-v1 dynamics (tests/cartpole.py): a
manual rollout with the same seeds must equal what BatchSampler returns -- nested lists and the packed store alike --,
across ten continuous calls too; and the vectorised sampler must reproduce E single-environment rollouts when the policy
is deterministic (its random streams differ by construction otherwise)."""
import numpy as np
from numpy.testing import assert_array_equal

from cartpole import CartPole
from rl_replicas_b200.packing import pack_experience
from rl_replicas_b200.policies import RandomPolicy
from rl_replicas_b200.samplers import BatchSampler, VectorSampler


def _manual_rollout(seed, num_samples):
    env = CartPole()
    env.action_space.seed(seed)
    out = {k: [] for k in ("observations", "actions", "rewards", "last_observations", "next_observations", "dones")}
    observation, _ = env.reset(seed=seed)
    for step in range(num_samples):
        out["observations"].append(observation)
        action = env.action_space.sample()
        out["actions"].append(action)
        observation, reward, terminated, truncated, _ = env.step(action)
        done = terminated or truncated
        out["next_observations"].append(observation)
        out["rewards"].append(reward)
        out["dones"].append(done)
        if done or step == num_samples - 1:
            out["last_observations"].append(observation)
            if done:
                observation, _ = env.reset()
    return out


def _sampler(seed, **kw):
    env = CartPole()
    env.action_space.seed(seed)
    return BatchSampler(env, seed, **kw), RandomPolicy(env.action_space)


def test_batch_sampler_equals_the_manual_rollout_on_cartpole():
    want = _manual_rollout(0, 1000)
    assert sum(want["dones"]) > 20  # random actions: many short episodes, so episode boundaries are exercised
    for packed in (False, True):
        sampler, policy = _sampler(0, packed=packed)
        exp = sampler.sample(1000, policy)
        assert_array_equal(exp.flattened_observations, want["observations"])
        assert_array_equal(np.asarray(exp.flattened_actions).reshape(-1), np.asarray(want["actions"]))
        assert_array_equal(exp.flattened_rewards, want["rewards"])
        assert_array_equal(exp.last_observations, want["last_observations"])
        assert_array_equal(exp.flattened_next_observations, want["next_observations"])
        assert_array_equal(exp.flattened_dones, want["dones"])
        b = pack_experience(exp)
        assert b["act"].shape == (1000,) and b["obs"].shape == (1000, 4)  # the categorical engine layout


def test_continuous_sampling_on_cartpole():
    want = _manual_rollout(0, 1000)
    sampler, policy = _sampler(0, is_continuous=True)
    got = {k: [] for k in ("observations", "actions", "rewards", "next_observations", "dones")}
    for _ in range(10):
        exp = sampler.sample(100, policy)
        got["observations"] += exp.flattened_observations
        got["actions"] += exp.flattened_actions
        got["rewards"] += exp.flattened_rewards
        got["next_observations"] += exp.flattened_next_observations
        got["dones"] += exp.flattened_dones
    for k in ("observations", "actions", "rewards", "next_observations"):
        assert_array_equal(got[k], want[k])
    # the cut-off at the end of a 100-step call closes an episode without ending it (ref batch_sampler.py:65-88):
    # every other flag agrees with the uninterrupted rollout
    assert_array_equal(got["dones"], want["dones"])


def test_vector_sampler_on_cartpole_matches_single_environment_rollouts():
    class Bang:  # deterministic, batch-transparent: push towards the side the pole leans to
        def get_action_numpy(self, observation):
            o = np.asarray(observation)
            return (o[..., 2] + 0.5 * o[..., 3] > 0).astype(np.int64)

    n_env, per_env = 4, 300
    vec_envs = [CartPole() for _ in range(n_env)]
    vec = VectorSampler(vec_envs, seed=10).sample(n_env * per_env, Bang())
    packed = pack_experience(vec)
    assert packed["act"].shape == (n_env * per_env,) and packed["ep_offsets"][-1] == n_env * per_env
    row = 0
    for e in range(n_env):
        single = pack_experience(BatchSampler(CartPole(), seed=10 + e).sample(per_env, Bang()))
        n = single["obs"].shape[0]
        assert_array_equal(packed["obs"][row:row + n], single["obs"])
        assert_array_equal(packed["act"][row:row + n], single["act"])
        assert_array_equal(packed["rew"][row:row + n], single["rew"])
        row += n
    assert row == n_env * per_env
