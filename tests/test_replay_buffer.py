"""Contiguous ReplayBuffer against a list-based restatement of the reference's FIFO semantics
(ref: replay_buffer.py:26-74: extend, delete from the front beyond buffer_size, np.random.randint + gather)."""
import numpy as np

from rl_replicas_b200.experience import Experience
from rl_replicas_b200.replay_buffer import ReplayBuffer


class _ListModel:
    def __init__(self, buffer_size):
        self.buffer_size, self.cols = buffer_size, [[], [], [], [], []]

    def add(self, exp):
        new = (exp.flattened_observations, exp.flattened_actions, exp.flattened_rewards,
               exp.flattened_next_observations, exp.flattened_dones)
        for c, v in zip(self.cols, new):
            c.extend(v)
        over = len(self.cols[0]) - self.buffer_size
        if over > 0:
            for c in self.cols:
                del c[:over]

    def sample(self, b):
        idx = np.random.randint(0, len(self.cols[0]), b)
        return [np.asarray([c[i] for i in idx]) for c in self.cols]


def _episode(rng, length, o=3, a=2):
    obs = [rng.standard_normal(o).astype(np.float32) for _ in range(length)]
    return dict(obs=obs, act=[rng.uniform(-1, 1, a).astype(np.float32) for _ in range(length)],
                rew=[float(rng.standard_normal()) for _ in range(length)], last=rng.standard_normal(o).astype(np.float32),
                done=[False] * (length - 1) + [bool(rng.integers(0, 2))])


def _experience(rng, lengths):
    eps = [_episode(rng, n) for n in lengths]
    return Experience([e["obs"] for e in eps], [e["act"] for e in eps], [e["rew"] for e in eps],
                      [e["last"] for e in eps], [e["done"] for e in eps])


def test_ring_buffer_matches_list_semantics_through_growth_and_wraparound():
    rng = np.random.default_rng(0)
    rb, model = ReplayBuffer(buffer_size=2500), _ListModel(2500)
    for round_ in range(12):
        exp = _experience(rng, rng.integers(1, 300, size=3))
        rb.add_experience(exp)
        model.add(exp)
        assert rb.current_size == len(model.cols[0]) <= 2500
        np.random.seed(100 + round_)
        got = rb.sample_minibatch(64)
        np.random.seed(100 + round_)
        want = model.sample(64)
        for k, w in zip(ReplayBuffer.COLUMNS, want):
            np.testing.assert_array_equal(np.asarray(got[k], dtype=w.dtype), w)
    # logical-order views (the reference's list attributes)
    np.testing.assert_array_equal(np.asarray(rb.observations), np.asarray(model.cols[0]))
    assert rb.rewards == model.cols[2] and rb.dones == model.cols[4]
    assert got["rewards"].dtype == np.float64 and got["dones"].dtype == np.bool_


def test_one_add_larger_than_the_buffer_keeps_the_newest():
    rng = np.random.default_rng(1)
    rb, model = ReplayBuffer(buffer_size=1100), _ListModel(1100)
    exp = _experience(rng, [700, 900])
    rb.add_experience(exp)
    model.add(exp)
    assert rb.current_size == 1100
    np.testing.assert_array_equal(np.asarray(rb.actions), np.asarray(model.cols[1]))


def test_batched_index_draws_consume_the_numpy_stream_like_repeated_sample_minibatch():
    rng = np.random.default_rng(2)
    rb = ReplayBuffer()
    rb.add_experience(_experience(rng, [50, 60]))
    np.random.seed(7)
    a = [rb.sample_minibatch(16) for _ in range(5)]
    np.random.seed(7)
    idx = np.stack([rb.sample_indices(16) for _ in range(5)])
    cols = rb.gather(idx)
    for s in range(5):
        for k in ReplayBuffer.COLUMNS:
            np.testing.assert_array_equal(cols[k][s], a[s][k])


def test_packed_experience_appends_by_columns_like_the_list_path():
    """ReplayBuffer.add_experience takes PackedExperience.transition_columns() (slice copies) instead of walking the
    flattened_* lists; both must leave the same transitions behind (ref: replay_buffer.py:30-49, experience.py:60-84)."""
    from rl_replicas_b200.experience import Experience, PackedExperience
    rng = np.random.default_rng(3)
    pe = PackedExperience(50, 4, 2)
    o = rng.standard_normal((51, 4)).astype(np.float32)
    for a, b in ((0, 20), (20, 50)):
        d = np.zeros(b - a, dtype=bool)
        d[-1] = a == 0
        pe.append_episode(o[a:b], rng.uniform(-1, 1, (b - a, 2)).astype(np.float32), rng.standard_normal(b - a), d, o[b])
    r1, r2 = ReplayBuffer(), ReplayBuffer()
    r1.add_experience(pe)
    r2.add_experience(Experience(observations=pe.observations, actions=pe.actions, rewards=pe.rewards,
                                 last_observations=pe.last_observations, dones=pe.dones))
    for k in ReplayBuffer.COLUMNS:
        np.testing.assert_array_equal(np.asarray(getattr(r1, k)), np.asarray(getattr(r2, k)))
    np.testing.assert_array_equal(np.asarray(r1.next_observations)[19], o[20])  # episode end -> its last observation
