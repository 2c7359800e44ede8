"""Pin the numpy oracle against outputs of the reference itself (tests/golden/*.npz).

CPU only.  Tolerances: 1e-5 of max|ref| per tensor (BASELINE.json north_star: 1e-5 relative fp32,
read norm-wise per SURVEY.md Appendix D), looser documented bounds for quantities that pass through
80 Adam steps (Adam's m/(sqrt(v)+eps) amplifies 1-ulp gradient differences of near-zero entries).
"""
import numpy as np
import pytest

from conftest import batch_of, load_golden, rel_err
from oracle import onpolicy as O

TOL = 1e-5


def test_scan_kats():
    g = load_golden("scan_kats")
    for fn in (O.discounted_cumulative_sums, O.discounted_cumulative_sums_lfilter):
        np.testing.assert_allclose(fn(g["dcs_in"], 0.5), g["dcs_out"], rtol=1e-15)
        np.testing.assert_allclose(fn(g["dcs_rand_in"], 0.99 * 0.97), g["dcs_rand_out"], rtol=1e-12, atol=1e-14)
    assert list(g["dcs_out"]) == [2.75, 3.5, 3.0]
    r, v = g["kat_r"], g["kat_v"]
    off = np.asarray([0, 3])
    for tag, done in (("done", True), ("notdone", False)):
        adv, ret = O.gae_and_returns(r, v[:3], v[3:], off, np.asarray([done]), 0.99, 0.97, fast=False)
        np.testing.assert_allclose(adv, g["gae_" + tag].astype(np.float32), rtol=1e-7)
        np.testing.assert_allclose(ret, g["ret_" + tag].astype(np.float32), rtol=1e-7)
    # reference quirk: advantages identical for done / not done
    np.testing.assert_array_equal(g["gae_done"], g["gae_notdone"])


CASES = ["ppo_categorical_cfg1", "ppo_gaussian_small", "ppo_gaussian_ragged_earlystop"]


def _setup(g):
    ps, vs = [int(x) for x in g["policy_sizes"]], [int(x) for x in g["value_sizes"]]
    policy = O.unflatten_layers(g["policy_flat0"], ps)
    value = O.unflatten_layers(g["value_flat0"], vs)
    kind = "gaussian" if "log_std" in g else "categorical"
    return policy, value, kind, g.get("log_std")


@pytest.mark.parametrize("case", CASES)
def test_preamble_matches_reference(case):
    g = load_golden(case)
    policy, value, kind, log_std = _setup(g)
    b = batch_of(g)
    values = O.mlp_forward(value, b["obs"])[0][:, 0]
    last_values = O.mlp_forward(value, b["last_obs"])[0][:, 0]
    assert rel_err(values, g["values"]) < TOL
    assert rel_err(last_values, g["last_values"]) < TOL
    # feed the reference's own values so the scan is compared in isolation: must be (near) bit-exact
    for fast in (True, False):
        adv_raw, ret = O.gae_and_returns(b["rew"], g["values"], g["last_values"], b["ep_offsets"], b["ep_done"],
                                         0.99, 0.97, fast=fast)
        assert rel_err(ret, g["ret"]) < 1e-7
        assert rel_err(adv_raw, g["adv_raw"]) < 1e-7
    assert rel_err(O.normalize(g["adv_raw"]), g["adv"]) < TOL
    old_logp = O.Dist(kind, O.mlp_forward(policy, b["obs"])[0], log_std).log_prob(b["act"])
    assert rel_err(old_logp, g["old_logp"]) < TOL


@pytest.mark.parametrize("case", CASES)
def test_first_step_gradients_and_adam(case):
    g = load_golden(case)
    policy, value, kind, log_std = _setup(g)
    b = batch_of(g)
    r = O.policy_loss_and_grad(policy, kind, log_std, b["obs"], b["act"], g["adv"], g["old_logp"], "ppo", 0.2)
    assert rel_err(r["grad"], g["grad0"]) < TOL
    assert abs(r["loss"] - g["metric:policy/loss"]) < 1e-6
    adam = O.AdamState(r["grad"].size, 3e-4)
    assert rel_err(adam.apply(g["policy_flat0"], g["grad0"]), g["policy_flat1"]) < 1e-6
    rv = O.value_loss_and_grad(value, b["obs"], g["ret"])
    assert rel_err(rv["grad"], g["vgrad0"]) < TOL
    assert abs(rv["loss"] - g["value_losses"][0]) < 1e-5 * abs(g["value_losses"][0])
    adam = O.AdamState(rv["grad"].size, 1e-3)
    assert rel_err(adam.apply(g["value_flat0"], g["vgrad0"]), g["value_flat1"]) < 1e-6


@pytest.mark.parametrize("case", CASES)
def test_full_ppo_train_matches_reference(case):
    g = load_golden(case)
    policy, value, kind, log_std = _setup(g)
    hp = dict(max_kl=float("inf")) if "inf" in str(g["hp_json"]) else {}
    out = O.ppo_train(batch_of(g), policy, value, kind, log_std, O.AdamState(g["policy_flat0"].size, 3e-4),
                      O.AdamState(g["value_flat0"].size, 1e-3), **hp)
    assert out["policy_steps"] == len(g["kl_trace"])  # same early-stop iteration
    assert rel_err(out["adv"], g["adv"]) < TOL
    assert rel_err(out["ret"], g["ret"]) < TOL
    # Value net: 80 Adam steps of a smooth loss -> stays tight.
    assert rel_err(out["value_flat"], g["value_flat_final"]) < 2e-5
    # Policy net: the PPO-clip gradient is DISCONTINUOUS in the parameters (a sample whose ratio sits at 1+-clip
    # flips its mask on a 1-ulp difference and changes the gradient by O(1/N)).  The trajectory therefore tracks the
    # reference tightly until the first flip (checked on the KL-trace prefix) and only loosely afterwards.
    n_prefix = min(30, len(g["kl_trace"]))
    assert rel_err(out["kl_trace"][:n_prefix], g["kl_trace"][:n_prefix]) < 1e-4
    assert rel_err(out["policy_flat"], g["policy_flat_final"]) < 1e-2
    assert rel_err(out["kl_trace"], g["kl_trace"]) < 2e-2
    assert rel_err(out["value_losses"], g["value_losses"]) < 1e-4
    assert abs(out["entropy_before"] - g["metric:policy/avarage_entropy"]) < 1e-5
    assert abs(out["logp_std_before"] - g["metric:policy/log_prob_std"]) < 1e-4 * g["metric:policy/log_prob_std"]
    assert abs(out["kl"] - g["metric:policy/kl_divergence"]) < 2e-2 * abs(g["metric:policy/kl_divergence"]) + 1e-7
    assert abs(out["value_loss_mean"] - g["metric:value_function/average_loss"]) < 1e-4 * g["metric:value_function/average_loss"]


def test_vpg_first_gradient():
    g = load_golden("vpg_gaussian_small")
    policy, value, kind, log_std = _setup(g)
    b = batch_of(g)
    values = O.mlp_forward(value, b["obs"])[0][:, 0]
    last_values = O.mlp_forward(value, b["last_obs"])[0][:, 0]
    adv_raw, ret = O.gae_and_returns(b["rew"], values, last_values, b["ep_offsets"], b["ep_done"], 0.99, 0.97)
    adv = O.normalize(adv_raw)
    r = O.policy_loss_and_grad(policy, kind, log_std, b["obs"], b["act"], adv, None, "vpg")
    assert rel_err(r["grad"], g["grad0"]) < TOL
    assert abs(r["loss"] - g["metric:policy/loss"]) < 1e-6


@pytest.mark.parametrize("case", CASES)
def test_torch_port_matches_reference(case):
    """oracle/torch_port.py (the CPU-baseline arm) issues the reference's own torch calls: it must track the
    reference's outputs to float32 round-off over the whole update."""
    import torch
    from oracle import torch_port as T
    g = load_golden(case)
    policy, value, kind, log_std = _setup(g)
    hp = dict(max_kl=float("inf")) if "inf" in str(g["hp_json"]) else {}
    nt = torch.get_num_threads()
    torch.set_num_threads(1)  # the fixtures were generated single-threaded
    try:
        out = T.ppo_train(batch_of(g), policy, value, kind, log_std, **hp)
    finally:
        torch.set_num_threads(nt)
    assert out["policy_steps"] == len(g["kl_trace"])
    assert rel_err(out["policy_flat"], g["policy_flat_final"]) < 1e-5
    assert rel_err(out["value_flat"], g["value_flat_final"]) < 1e-5
    assert rel_err(out["value_losses"], g["value_losses"]) < 1e-5
    assert rel_err(out["ret"], g["ret"]) < 1e-6 and rel_err(out["adv"], g["adv"]) < 1e-5


TRPO_CASES = ["trpo_gaussian_small", "trpo_categorical_small"]


@pytest.mark.parametrize("case", TRPO_CASES)
def test_trpo_pieces_match_reference(case):
    """Analytic Fisher-vector product == the reference's double-backprop Hessian-vector product at theta_old; CG
    solution, descent step and the line-search outcome match the reference's ConjugateGradientOptimizer."""
    g = load_golden(case)
    policy, value, kind, log_std = _setup(g)
    b = batch_of(g)
    hv = O.fisher_vector_product(policy, kind, log_std, b["obs"], g["hvp_probe"], 1e-5)
    assert rel_err(hv, g["hvp_of_probe"]) < 1e-5
    out = O.trpo_policy_step(policy, kind, log_std, b["obs"], b["act"], g["adv"])
    assert rel_err(out["grad"], g["grad0"]) < TOL
    # CG amplifies round-off over 10 iterations: documented looser bounds
    assert rel_err(out["x"], g["cg_x"]) < 2e-3
    assert rel_err(out["descent"], g["descent"]) < 2e-3
    assert rel_err(out["policy_flat"], g["policy_flat_final"]) < 2e-3
    assert not out["rejected"]
    assert abs(out["kl"] - g["final_kl"]) < 2e-2 * g["final_kl"]
    assert abs(out["new_loss"] - g["final_loss"]) < 2e-2 * abs(g["final_loss"])
    assert abs(out["loss_before"] - g["metric:policy/loss"]) < 1e-6
