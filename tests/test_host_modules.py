"""Host-side API objects the update engine reads its state from (no GPU): the MLP module tree and initialisation,
the critic wrappers, the metrics manager (ref: networks/mlp.py:15-41, value_function.py:5-28, q_function.py:6-32,
metrics_manager.py:11-42)."""
import os

import pytest
import torch
from torch import nn

from rl_replicas_b200.metrics_manager import MetricsManager
from rl_replicas_b200.networks import MLP
from rl_replicas_b200.q_function import QFunction
from rl_replicas_b200.value_function import ValueFunction


def test_mlp_has_the_reference_module_tree_and_initialisation():
    torch.manual_seed(3)
    net = MLP([5, 64, 32, 2], nn.ReLU, nn.Tanh)
    torch.manual_seed(3)  # the reference builds Linear, act, Linear, act, Linear, out-act in this order
    expected = nn.Sequential(nn.Linear(5, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 2), nn.Tanh())
    assert list(net.state_dict()) == ["network." + k for k in expected.state_dict()]  # checkpoint keys
    for ours, theirs in zip(net.state_dict().values(), expected.state_dict().values()):
        assert torch.equal(ours, theirs)
    assert [type(m) for m in net.network] == [type(m) for m in expected]
    x = torch.randn(7, 5)
    assert torch.equal(net(x), expected(x))
    with pytest.raises(ValueError):
        MLP([4])


def test_critics_forward_and_expose_network_and_optimizer():
    vnet, qnet = MLP([3, 8, 8, 1]), MLP([5, 8, 8, 1])
    v = ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3))
    q = QFunction(qnet, torch.optim.Adam(qnet.parameters(), lr=1e-3))
    obs, act = torch.randn(6, 3), torch.randn(6, 2)
    assert v(obs).shape == (6, 1) and torch.equal(v(obs), vnet(obs))
    assert q(obs, act).shape == (6,) and torch.equal(q(obs, act), qnet(torch.cat([obs, act], -1)).squeeze(-1))
    assert v.network is vnet and isinstance(q.optimizer, torch.optim.Adam)
    assert sorted(k for k, _ in v.named_parameters()) == sorted("network." + k for k, _ in vnet.named_parameters())


def test_metrics_manager_prints_every_scalar_and_keeps_tensorboard_optional(tmp_path, capsys):
    mm = MetricsManager(str(tmp_path))
    mm.record_scalar("epoch", 3)
    mm.record_scalar("policy/loss", -0.0123456, 4000, tensorboard=True)
    mm.dump()
    mm.close()
    out = capsys.readouterr().out.splitlines()
    assert out == ["epoch: 3       ", "policy/loss: -0.0123 "]  # the reference's "{}: {:<8.3g}" lines
    if mm.tensorboard_writer is not None:  # TensorBoard installed: an event file under <log_dir>/tensorboard
        assert any(f.startswith("events.out.tfevents") for f in os.listdir(os.path.join(tmp_path, "tensorboard")))


def test_synthetic_learners_hold_the_given_parameters():
    """bench.py builds its learners from plain (W, b) pairs through rl_replicas_b200.synthetic (no oracle import on the
    measured arm): the flat order is torch's parameters() order and the recipes' optimizers are attached."""
    import numpy as np

    from rl_replicas_b200 import synthetic
    from rl_replicas_b200.optimizers import ConjugateGradientOptimizer
    rng = np.random.default_rng(0)
    mk = lambda sz: [(rng.standard_normal((o, i)).astype(np.float32), rng.standard_normal(o).astype(np.float32))
                     for i, o in zip(sz[:-1], sz[1:])]
    pl, vl = mk([5, 64, 32, 3]), mk([5, 64, 32, 1])
    vec = lambda m: torch.nn.utils.parameters_to_vector(m.parameters()).detach().numpy()
    ppo = synthetic.onpolicy_learner("ppo", pl, vl, np.full(3, -0.5, np.float32), num_policy_gradients=2)
    assert np.array_equal(vec(ppo.policy.network), synthetic.flatten_layers(pl))
    assert np.array_equal(vec(ppo.value_function.network), synthetic.flatten_layers(vl))
    assert ppo.num_policy_gradients == 2 and ppo.policy.optimizer.defaults["lr"] == 3e-4
    assert np.array_equal(vec(ppo.old_policy.network), synthetic.flatten_layers(pl))
    trpo = synthetic.onpolicy_learner("trpo", pl, vl)  # no log_std: categorical
    assert isinstance(trpo.policy.optimizer, ConjugateGradientOptimizer) and type(trpo.policy).__name__ == "CategoricalPolicy"
    x = rng.standard_normal((4, 5)).astype(np.float32)
    np.testing.assert_allclose(synthetic.numpy_mlp(pl, x), ppo.policy.network(torch.from_numpy(x)).detach().numpy(),
                               rtol=1e-5, atol=1e-5)
    td3, buffer = synthetic.offpolicy_learner(True, mk([4, 16, 16, 2]), [mk([6, 16, 16, 1]), mk([6, 16, 16, 1])])
    assert type(td3).__name__ == "TD3" and td3.replay_buffer is buffer
    assert np.array_equal(vec(td3.target_policy.network), vec(td3.policy.network))
    ddpg, _ = synthetic.offpolicy_learner(False, mk([4, 16, 16, 2]), [mk([6, 16, 16, 1])])
    assert type(ddpg).__name__ == "DDPG"
