"""CPU: the numpy oracle against the round-2 goldens (reference outputs, tests/golden/make_golden_r2.py): the
ConjugateGradientOptimizer corner cases by value, the float64 scan helpers, and that the host-side sampling path
(torch CPU ops, policies/base.py) reproduces the reference's draws bit for bit from the same parameters."""
import numpy as np
import pytest
import torch

from conftest import batch_of, load_golden, rel_err
from oracle import onpolicy as O


def _trpo_case(name):
    g = load_golden(name)
    damping, delta, backtracks, lam = [float(x) for x in g["hp"]]
    policy = O.unflatten_layers(g["policy_flat0"], [27, 64, 64, 8])
    value = O.unflatten_layers(g["value_flat0"], [27, 64, 64, 1])
    b = batch_of(g)
    values = O.mlp_forward(value, b["obs"])[0][:, 0]
    last_values = O.mlp_forward(value, b["last_obs"])[0][:, 0]
    adv_raw, _ = O.gae_and_returns(b["rew"], values, last_values, b["ep_offsets"], b["ep_done"], 0.99, lam)
    with np.errstate(all="ignore"):
        adv = O.normalize(adv_raw)
        out = O.trpo_policy_step(policy, "gaussian", g["log_std"], b["obs"], b["act"], adv, max_constraint=delta,
                                 max_backtracks=int(backtracks), damping=damping)
    return g, out


def test_oracle_nan_step_size_rule_and_rejection():
    g, out = _trpo_case("trpo_reject_negdamp")
    assert out["xhx"] < 0 and out["step_size"] == 1.0 and out["rejected"] and out["accepted"] == -1
    assert rel_err(out["descent"], g["descent"]) < 2e-3
    np.testing.assert_array_equal(out["policy_flat"], g["policy_flat0"])
    assert abs(out["new_loss"] - g["ls_loss"][-1]) < 2e-2 * abs(g["ls_loss"][-1])


def test_oracle_oversized_trust_region_is_rejected():
    g, out = _trpo_case("trpo_reject_bigdelta")
    assert out["rejected"] and out["accepted"] == -1
    assert rel_err(out["descent"], g["descent"]) < 2e-3
    assert abs(out["kl"] - g["ls_kl"][-1]) < 5e-2 * g["ls_kl"][-1]
    np.testing.assert_array_equal(out["policy_flat"], g["policy_flat0"])


def test_oracle_nan_direction_rule():
    g, out = _trpo_case("trpo_reject_nan_adv")
    assert np.all(out["x"] == 0) and np.all(out["descent"] == 0) and out["rejected"] and np.isnan(out["new_loss"])
    np.testing.assert_array_equal(out["policy_flat"], g["policy_flat0"])


def test_oracle_scan_helpers_match_the_reference_utils():
    u = load_golden("utils_kats")
    for c in range(3):
        np.testing.assert_allclose(O.discounted_cumulative_sums(u["dcs2d_in"][:, c], 0.9), u["dcs2d_out"][:, c], rtol=1e-12,
                                   atol=1e-13)
    np.testing.assert_allclose(O.discounted_cumulative_sums(u["dcs_long_in"], 0.999), u["dcs_long_out"], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("case", ["sampled_actions_categorical", "sampled_actions_gaussian"])
def test_host_sampling_reproduces_the_reference_draws(case):
    """policies/base.py samples with the reference's torch CPU calls: from the reference's final parameters the draws
    are bit-identical (indices AND Gaussian samples); the GPU test repeats this from the parameters a GPU update leaves."""
    from rl_replicas_b200.algorithms._onpolicy import describe_mlp, write_flat
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.policies import CategoricalPolicy, GaussianPolicy
    g = load_golden(case)
    discrete = "log_std" not in g
    net = MLP([4, 64, 64, 3] if discrete else [17, 64, 64, 6])
    write_flat(describe_mlp(net)[3], g["policy_flat_final"])
    opt = torch.optim.Adam(net.parameters(), lr=3e-4)
    policy = (CategoricalPolicy(net, opt) if discrete else
              GaussianPolicy(net, opt, torch.nn.Parameter(torch.from_numpy(g["log_std"]))))
    torch.manual_seed(1234)
    single = np.stack([np.asarray(policy.get_action_numpy(g["probe_obs"][i])) for i in range(1000)])
    batched = np.asarray(policy.get_action_numpy(g["probe_obs"]))
    np.testing.assert_array_equal(single, g["single_draws"])
    np.testing.assert_array_equal(batched, g["batched_draw"])


@pytest.mark.parametrize("case", ["trpo_gaussian_small", "trpo_reject_negdamp", "trpo_reject_nan_adv"])
def test_cg_optimizer_step_with_user_torch_closures_matches_reference(case):
    """ConjugateGradientOptimizer.step(loss_fn, kl_fn) on plain torch closures (a caller other than TRPO.train): the
    torch-tensor route of optimizers/conjugate_gradient_optimizer.py, against what the reference's optimizer did on
    the same closures (ref: conjugate_gradient_optimizer.py:59-98; closures as in trpo.py:154-175)."""
    from torch.distributions import kl
    from rl_replicas_b200.algorithms._onpolicy import describe_mlp, write_flat
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.optimizers import ConjugateGradientOptimizer
    from rl_replicas_b200.policies import GaussianPolicy
    g = load_golden(case)
    damping, delta, backtracks, lam = [float(x) for x in g["hp"]] if "hp" in g else (1e-5, 0.01, 15, 0.97)
    net, old_net = MLP([27, 64, 64, 8]), MLP([27, 64, 64, 8])
    for n in (net, old_net):
        write_flat(describe_mlp(n)[3], g["policy_flat0"])
    opt = ConjugateGradientOptimizer(net.parameters(), max_constraint=delta, hvp_damping_coefficient=damping,
                                     max_backtracks=int(backtracks))
    log_std = torch.nn.Parameter(torch.from_numpy(g["log_std"]))
    policy, old_policy = GaussianPolicy(net, opt, log_std), GaussianPolicy(old_net, None, log_std)
    obs, act = torch.from_numpy(g["obs"]), torch.from_numpy(g["act"])
    if "adv" in g:
        adv = torch.from_numpy(g["adv"])
    else:  # the reject goldens store no advantages: rebuild them with the oracle's scan
        value = O.unflatten_layers(g["value_flat0"], [27, 64, 64, 1])
        b = batch_of(g)
        adv_raw, _ = O.gae_and_returns(b["rew"], O.mlp_forward(value, b["obs"])[0][:, 0],
                                       O.mlp_forward(value, b["last_obs"])[0][:, 0], b["ep_offsets"], b["ep_done"], 0.99, lam)
        t = torch.from_numpy(adv_raw.astype(np.float32))
        adv = (t - t.mean()) / t.std()

    def loss_fn():
        with torch.no_grad():
            old = old_policy(obs).log_prob(act)
        return -torch.mean(torch.exp(policy(obs).log_prob(act) - old) * adv)

    def kl_fn():
        with torch.no_grad():
            old = old_policy(obs)
        return torch.mean(kl.kl_divergence(old, policy(obs)))

    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        loss = loss_fn()
        opt.zero_grad()
        loss.backward()
        opt.step(loss_fn, kl_fn)
    finally:
        torch.set_num_threads(nt)
    got = torch.nn.utils.parameters_to_vector(net.parameters()).detach().numpy()
    if case == "trpo_gaussian_small":
        assert rel_err(got, g["policy_flat_final"]) < 1e-5
        assert np.abs(got - g["policy_flat0"]).max() > 1e-4  # the step was taken
    else:
        np.testing.assert_array_equal(got, g["policy_flat0"])  # rejected and restored, as in the reference


def test_native_closures_refuse_to_be_called_outside_the_optimizer():
    from rl_replicas_b200.optimizers.conjugate_gradient_optimizer import NativeClosure
    with pytest.raises(NotImplementedError):
        NativeClosure("surrogate loss", lambda opt: None)()


def test_oracle_trains_log_std_like_the_reference():
    """log_std inside the policy optimizer (policies/gaussian_policy.py:25-37): six PPO steps move it; the oracle's
    closed-form d logp / d log_std = z^2 - 1 must follow the reference's autograd."""
    g = load_golden("ppo_trainable_log_std")
    policy = O.unflatten_layers(g["policy_flat0"], [17, 64, 64, 6])
    value = O.unflatten_layers(g["value_flat0"], [17, 64, 64, 1])
    out = O.ppo_train(batch_of(g), policy, value, "gaussian", g["log_std0"], O.AdamState(5702 + 6, 3e-4),
                      O.AdamState(5377, 1e-3), max_kl=float("inf"), n_policy=6, n_value=3, train_log_std=True)
    assert np.abs(g["log_std_final"] - g["log_std0"]).max() > 1e-3  # it really moved
    assert rel_err(out["log_std"], g["log_std_final"]) < 1e-5
    assert rel_err(out["policy_flat"], g["policy_flat_final"]) < 1e-5
    assert rel_err(out["value_flat"], g["value_flat_final"]) < 1e-5
    assert abs(out["kl"] - g["metric:policy/kl_divergence"]) < 1e-4 * abs(g["metric:policy/kl_divergence"]) + 1e-8
