"""Trainable log_std (ref policies/gaussian_policy.py:25-37 with log_std inside the policy optimizer): the GPU update
differentiates through it and Adam moves it, as the reference's autograd does (golden generated from the reference)."""
import numpy as np
import pytest
import torch

from conftest import batch_of, load_golden, rel_err
from test_gpu_ppo import Rec, flat

pytestmark = pytest.mark.gpu


def _build(g, algo_cls, **hp):
    from rl_replicas_b200.algorithms._onpolicy import describe_mlp, write_flat
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.policies import GaussianPolicy
    from rl_replicas_b200.value_function import ValueFunction
    pnet, vnet = MLP([17, 64, 64, 6]), MLP([17, 64, 64, 1])
    write_flat(describe_mlp(pnet)[3], g["policy_flat0"])
    write_flat(describe_mlp(vnet)[3], g["value_flat0"])
    log_std = torch.nn.Parameter(torch.from_numpy(g["log_std0"].copy()))
    policy = GaussianPolicy(pnet, torch.optim.Adam(list(pnet.parameters()) + [log_std], lr=3e-4), log_std)
    vf = ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3))
    algo = algo_cls(policy, vf, None, None, **hp)
    algo.metrics_manager = Rec()
    algo.current_total_steps = 0
    return algo


def test_ppo_trains_log_std_like_the_reference():
    from rl_replicas_b200.algorithms import PPO
    g = load_golden("ppo_trainable_log_std")
    ppo = _build(g, PPO, num_policy_gradients=6, num_value_gradients=3, max_kl_divergence=float("inf"))
    ppo.train_packed(batch_of(g))
    assert ppo._engine.train_log_std and ppo.last_update_stats.fused == 0
    got = ppo.policy.log_std.detach().numpy()
    assert np.abs(got - g["log_std0"]).max() > 1e-3
    assert rel_err(got, g["log_std_final"]) < 1e-5
    assert rel_err(flat(ppo.policy.network), g["policy_flat_final"]) < 1e-5
    assert rel_err(flat(ppo.value_function.network), g["value_flat_final"]) < 1e-5
    np.testing.assert_array_equal(ppo.old_policy.log_std.detach().numpy(), got)  # ppo.py:183 syncs the whole policy
    st = ppo.policy.optimizer.state[ppo.policy.log_std]
    assert float(st["step"]) == g["log_std_adam_step"]
    assert rel_err(st["exp_avg"].numpy(), g["log_std_adam_m"]) < 1e-4
    assert rel_err(st["exp_avg_sq"].numpy(), g["log_std_adam_v"]) < 1e-4
    m = ppo.metrics_manager.s
    assert abs(m["policy/kl_divergence"] - g["metric:policy/kl_divergence"]) < 1e-4 * abs(g["metric:policy/kl_divergence"]) + 1e-8
    # a second update continues from the trained value (state round trip through the host modules)
    before = got.copy()
    ppo.train_packed(batch_of(g))
    assert np.abs(ppo.policy.log_std.detach().numpy() - before).max() > 1e-4


def test_vpg_trains_log_std_like_the_oracle():
    from oracle import onpolicy as O
    from rl_replicas_b200.algorithms import VPG
    g = load_golden("ppo_trainable_log_std")
    vpg = _build(g, VPG, num_value_gradients=2)
    vpg.train_packed(batch_of(g))
    b = batch_of(g)
    policy = O.unflatten_layers(g["policy_flat0"], [17, 64, 64, 6])
    value = O.unflatten_layers(g["value_flat0"], [17, 64, 64, 1])
    values = O.mlp_forward(value, b["obs"])[0][:, 0]
    last_values = O.mlp_forward(value, b["last_obs"])[0][:, 0]
    adv = O.normalize(O.gae_and_returns(b["rew"], values, last_values, b["ep_offsets"], b["ep_done"], 0.99, 0.97)[0])
    r = O.policy_loss_and_grad(policy, "gaussian", g["log_std0"], b["obs"], b["act"], adv, None, "vpg")
    want = O.AdamState(5708, 3e-4).apply(np.concatenate([g["policy_flat0"], g["log_std0"]]),
                                         np.concatenate([r["grad"], r["grad_log_std"]]))
    assert rel_err(flat(vpg.policy.network), want[:5702]) < 1e-6
    assert rel_err(vpg.policy.log_std.detach().numpy(), want[5702:]) < 1e-6


def test_trpo_refuses_a_log_std_inside_the_cg_optimizer():
    from rl_replicas_b200.algorithms import TRPO
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.optimizers import ConjugateGradientOptimizer
    from rl_replicas_b200.policies import GaussianPolicy
    from rl_replicas_b200.value_function import ValueFunction
    g = load_golden("ppo_trainable_log_std")
    pnet, vnet = MLP([17, 64, 64, 6]), MLP([17, 64, 64, 1])
    log_std = torch.nn.Parameter(torch.from_numpy(g["log_std0"].copy()))
    policy = GaussianPolicy(pnet, ConjugateGradientOptimizer(list(pnet.parameters()) + [log_std]), log_std)
    trpo = TRPO(policy, ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3)), None, None)
    with pytest.raises(NotImplementedError):
        trpo.train_packed(batch_of(g))
