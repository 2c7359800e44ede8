"""Host-side logic of TD3 / DDPG's learner-state transfer (algorithms/td3.py: _state_plan, _upload_state,
_download_state) against a stand-in engine: blob layout (segments on multiples of 64 floats), round trip through the
host modules, Adam moments and step counts.  No GPU."""
import numpy as np
import pytest
import torch

from rl_replicas_b200 import synthetic
from rl_replicas_b200.algorithms._onpolicy import describe_mlp


class _BlobEngine:
    """state_layout / state_buffer / get_state / set_state of OffPolicyEngine over a plain host buffer."""

    def __init__(self, sizes, n_q):
        self.sizes, self.n_q, self.steps = sizes, n_q, [0, 0, 0]
        self.buf = torch.zeros(self.state_layout()[1])

    def state_layout(self):
        pad = lambda n: (n + 63) & ~63
        present = [0, 1] + ([2] if self.n_q == 2 else []) + [3, 4] + ([5] if self.n_q == 2 else [])
        out, off = [], 0
        for i in present:
            out.append(("params", i, off, self.sizes[i]))
            off += pad(self.sizes[i])
        for i in [0, 1] + ([2] if self.n_q == 2 else []):
            for kind in ("m", "v"):
                out.append((kind, i, off, self.sizes[i]))
                off += pad(self.sizes[i])
        return out, off

    def state_buffer(self):
        return self.buf

    def set_state(self, blob, steps):
        assert blob is None  # the algorithms fill state_buffer() in place
        self.steps = list(steps)

    def get_state(self):
        return self.buf.numpy(), list(self.steps)


@pytest.mark.parametrize("twin", [True, False])
def test_state_round_trip_through_the_blob(twin):
    rng = np.random.default_rng(7)
    mk = lambda sz: [(rng.standard_normal((o, i)).astype(np.float32), rng.standard_normal(o).astype(np.float32))
                     for i, o in zip(sz[:-1], sz[1:])]
    qs = [mk([7, 24, 12, 1]) for _ in range(2 if twin else 1)]
    algo, _ = synthetic.offpolicy_learner(twin, mk([5, 24, 12, 2]), qs)
    trainable, targets = algo._nets()
    count = lambda m: sum(p.numel() for p in m.network.parameters())
    sizes = {i: count(m) for i, m in enumerate(trainable)}
    sizes.update({3 + i: count(m) for i, m in enumerate(targets)})
    e = _BlobEngine(sizes, 2 if twin else 1)
    lins = [describe_mlp(m.network)[3] for m in trainable + targets]
    flat = lambda m: torch.nn.utils.parameters_to_vector(m.network.parameters()).detach().numpy().copy()

    algo._upload_state(e, trainable, targets, lins)
    layout, total = e.state_layout()
    assert total % 64 == 0 and e.steps == [0, 0, 0]
    blob = e.buf.numpy()
    mods = dict(enumerate(trainable))
    mods.update({3 + i: m for i, m in enumerate(targets)})
    for kind, i, off, n in layout:
        if kind == "params":
            np.testing.assert_array_equal(blob[off:off + n], flat(mods[i]))
        else:
            assert not blob[off:off + n].any()  # no Adam state yet: zeros

    # the engine "trains": new parameters, moments and step counts come back in the blob
    want = {}
    for kind, i, off, n in layout:
        x = rng.standard_normal(n).astype(np.float32)
        blob[off:off + n] = np.abs(x) if kind == "v" else x
        want[(kind, i)] = blob[off:off + n].copy()
    e.steps = [3, 5, 5 if twin else 0]
    algo._download_state(e, trainable, targets, lins)
    for i, m in mods.items():
        np.testing.assert_array_equal(flat(m), want[("params", i)])
    for i, m in enumerate(trainable):
        ps = list(m.network.parameters())
        st = m.optimizer.state
        np.testing.assert_array_equal(np.concatenate([st[p]["exp_avg"].numpy().ravel() for p in ps]), want[("m", i)])
        np.testing.assert_array_equal(np.concatenate([st[p]["exp_avg_sq"].numpy().ravel() for p in ps]), want[("v", i)])
        assert all(float(st[p]["step"]) == e.steps[i] for p in ps)

    # second upload: the moments and steps written above travel back unchanged, into the SAME cached plan
    plan = algo._plan
    blob[:] = np.nan
    algo._upload_state(e, trainable, targets, lins)
    assert algo._plan is plan and e.steps[:len(trainable)] == [3, 5, 5][:len(trainable)]
    for kind, i, off, n in layout:
        np.testing.assert_array_equal(blob[off:off + n], want[(kind, i)])
