"""PackedExperience: the contiguous rollout store gives the engine the same arrays as flattening the
reference's nested lists, and keeps the nested-list API as views."""
import numpy as np

from rl_replicas_b200.packing import pack_experience
from rl_replicas_b200.samplers import BatchSampler


class _Env:
    """Deterministic toy env (gymnasium protocol): episodes end after 7, 3, 12, ... steps (terminated) -- no physics."""

    def __init__(self):
        self.rng = np.random.default_rng(0)
        self.t, self.length = 0, 7

    def reset(self, seed=None):
        self.t = 0
        self.length = int(self.rng.integers(3, 13))
        return self.rng.standard_normal(4).astype(np.float32), {}

    def step(self, action):
        self.t += 1
        obs = (self.rng.standard_normal(4) + 0.1 * float(np.sum(action))).astype(np.float32)
        return obs, float(self.rng.standard_normal()), self.t >= self.length, False, {}


class _Policy:
    def __init__(self):
        self.rng = np.random.default_rng(5)

    def get_action_numpy(self, observation):
        return self.rng.uniform(-1, 1, 2).astype(np.float32)


def test_packed_rollout_equals_flattened_nested_lists():
    nested = BatchSampler(_Env(), seed=0).sample(100, _Policy())
    packed = BatchSampler(_Env(), seed=0, packed=True).sample(100, _Policy())
    a, b = pack_experience(nested), pack_experience(packed)
    assert set(a) == set(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
        assert a[k].dtype == b[k].dtype
    assert b["obs"].flags["C_CONTIGUOUS"] and b["obs"].base is not None  # a view of the backing store, not a copy
    # nested-list API
    assert packed.episode_lengths == nested.episode_lengths
    np.testing.assert_allclose(packed.episode_returns, nested.episode_returns)
    assert packed.episode_dones == nested.episode_dones
    for ep_p, ep_n in zip(packed.observations, nested.observations):
        np.testing.assert_array_equal(np.asarray(ep_p), np.asarray(ep_n))
    np.testing.assert_array_equal(np.asarray(packed.flattened_next_observations),
                                  np.asarray(nested.flattened_next_observations))
    assert packed.flattened_rewards == nested.flattened_rewards and packed.flattened_dones == nested.flattened_dones


def test_packed_experience_feeds_the_replay_buffer():
    from rl_replicas_b200.replay_buffer import ReplayBuffer
    nested = BatchSampler(_Env(), seed=0).sample(60, _Policy())
    packed = BatchSampler(_Env(), seed=0, packed=True).sample(60, _Policy())
    ra, rb = ReplayBuffer(), ReplayBuffer()
    ra.add_experience(nested)
    rb.add_experience(packed)
    np.random.seed(1)
    x = ra.sample_minibatch(32)
    np.random.seed(1)
    y = rb.sample_minibatch(32)
    for k in x:
        np.testing.assert_array_equal(x[k], y[k])


def test_vector_sampler_lockstep_rollouts_have_the_single_env_structure():
    """VectorSampler: E toy environments in lock step, one batched policy call per step.  With a
    deterministic policy each environment's rollout must equal what a BatchSampler produces on a twin environment."""
    from rl_replicas_b200.samplers import VectorSampler

    class Det:
        def get_action_numpy(self, observation):
            o = np.asarray(observation, np.float32)
            return np.tanh(o[..., :2] * 0.5).astype(np.float32)  # batch-transparent

    class SeededEnv(_Env):
        def __init__(self, seed):
            super().__init__()
            self.rng = np.random.default_rng(seed)

    n_env, per_env = 3, 40
    vec = VectorSampler([SeededEnv(10 + e) for e in range(n_env)]).sample(n_env * per_env, Det())
    packed = pack_experience(vec)
    assert packed["obs"].shape == (n_env * per_env, 4) and packed["ep_offsets"][-1] == n_env * per_env
    row = 0
    for e in range(n_env):
        single = pack_experience(BatchSampler(SeededEnv(10 + e)).sample(per_env, Det()))
        n = single["obs"].shape[0]
        np.testing.assert_array_equal(packed["obs"][row:row + n], single["obs"])
        np.testing.assert_array_equal(packed["act"][row:row + n], single["act"])
        np.testing.assert_array_equal(packed["rew"][row:row + n], single["rew"])
        row += n
    assert row == n_env * per_env
    assert sum(vec.episode_lengths) == n_env * per_env and len(vec.episode_dones) == len(vec.episode_lengths)


class _DiscretePolicy:
    def __init__(self):
        self.rng = np.random.default_rng(7)

    def get_action_numpy(self, observation):
        o = np.asarray(observation)
        if o.ndim == 2:  # batched call of the VectorSampler
            return self.rng.integers(0, 3, o.shape[0])
        return np.asarray(self.rng.integers(0, 3))  # 0-d, like CategoricalPolicy.get_action_numpy on one observation


def test_packed_rollout_with_scalar_actions_has_the_categorical_layout():
    """Discrete action spaces: the reference's lists hold 0-d actions and the engine's categorical path wants
    act [N] -- the packed store must produce exactly what flattening the nested lists produces."""
    from rl_replicas_b200.samplers import VectorSampler
    nested = BatchSampler(_Env(), seed=0).sample(50, _DiscretePolicy())
    packed = BatchSampler(_Env(), seed=0, packed=True).sample(50, _DiscretePolicy())
    a, b = pack_experience(nested), pack_experience(packed)
    assert a["act"].shape == (50,) and b["act"].shape == (50,)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert np.asarray(packed.actions[0][0]).ndim == 0
    vec = pack_experience(VectorSampler([_Env(), _Env()]).sample(40, _DiscretePolicy()))
    assert vec["act"].shape == (40,) and vec["act"].dtype == np.float32
