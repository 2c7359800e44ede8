"""Round-2 golden vectors, again produced by running the UNMODIFIED reference (rl_replicas @ /root/reference):

    python tests/golden/make_golden_r2.py        (build container only; the GPU box has no /root/reference)

  trpo_reject_*.npz   ConjugateGradientOptimizer corner cases taken by value from the reference
                      (optimizers/conjugate_gradient_optimizer.py:83 NaN direction -> 0, :92-93 NaN step size -> 1.0,
                      :233-250 reject / restore): a negative damping coefficient (x^T H x < 0 => sqrt of a negative
                      number), a trust region far too large for the quadratic model, and advantages that are all equal
                      (normalize_tensor divides 0 by 0).  Every case is rejected with a macroscopic margin.
  sampled_actions_*.npz  actions drawn by the reference's policies (policies/stochastic_policy.py:26-41) from the
                      parameters the reference's PPO.train leaves behind, under a fixed torch seed: the "bit-exact
                      sampled action indices" clause of BASELINE.json.
  utils_kats.npz      inputs / outputs of the reference's public helpers (utils.py:14-92).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (stubs gymnasium, imports the reference and rl_replicas_b200.synthetic)
from rl_replicas.algorithms import PPO, TRPO  # noqa: E402
from rl_replicas.experience import Experience  # noqa: E402
from rl_replicas.networks import MLP  # noqa: E402
from rl_replicas.optimizers import ConjugateGradientOptimizer  # noqa: E402
from rl_replicas.policies import GaussianPolicy  # noqa: E402
from rl_replicas.utils import (bootstrap_rewards_with_last_values, compute_values, discounted_cumulative_sums, gae,  # noqa: E402
                               normalize_tensor, polyak_average, set_seed_for_libraries)
from rl_replicas.value_function import ValueFunction  # noqa: E402

synthetic = G.synthetic


def trpo_reject_case(name, damping=1e-5, delta=0.01, backtracks=15, const_adv=False):
    set_seed_for_libraries(0)
    obs_dim, act_dim = 27, 8
    pnet, vnet = MLP([obs_dim, 64, 64, act_dim]), MLP([obs_dim, 64, 64, 1])
    opt = ConjugateGradientOptimizer(pnet.parameters(), max_constraint=delta, hvp_damping_coefficient=damping,
                                     max_backtracks=backtracks)
    log_std = torch.nn.Parameter(-0.5 * torch.ones(act_dim))
    policy = GaussianPolicy(pnet, opt, log_std)
    vf = ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3))
    with torch.no_grad():
        batch = synthetic.fixed_batch(8, 150, obs_dim, act_dim, seed=4, frac_not_done=0.3,
                                      mean_fn=lambda o: policy.network(torch.from_numpy(o)).numpy())
    gae_lambda = 0.97
    if const_adv:  # V == 0 and r == 1 everywhere, lambda = 0: every TD residual is exactly 1.0 => std == 0
        batch["rew"][:] = 1.0
        with torch.no_grad():
            for p in list(vnet.parameters())[-2:]:
                p.zero_()
        gae_lambda = 0.0
    exp = Experience(**synthetic.to_experience_lists(batch, False))
    trpo = TRPO(policy, vf, None, None, num_value_gradients=2, gae_lambda=gae_lambda)
    trpo.metrics_manager = G.Recorder()
    trpo.current_total_steps = 0
    out = dict(batch)
    out["policy_flat0"], out["value_flat0"] = G.flat(pnet), G.flat(vnet)
    out["log_std"] = log_std.detach().numpy().copy()
    cap = {"ls_loss": [], "ls_kl": []}
    orig_cg, orig_ls = opt._conjugate_gradient, opt._backtracking_line_search

    def cg_wrap(hvp, b, residual_tol=1e-10):
        cap["grad0"] = b.detach().numpy().copy()
        x = orig_cg(hvp, b, residual_tol)
        cap["cg_x_raw"] = x.detach().numpy().copy()  # before the NaN -> 0 rule
        return x

    def ls_wrap(params, descent_step, loss_fn, kl_fn):
        cap["descent"] = torch.as_tensor(descent_step).detach().numpy().copy()

        def lf():
            r = loss_fn()
            cap["ls_loss"].append(float(r))
            return r

        def kf():
            r = kl_fn()
            cap["ls_kl"].append(float(r))
            return r

        orig_ls(params, descent_step, lf, kf)

    opt._conjugate_gradient, opt._backtracking_line_search = cg_wrap, ls_wrap
    trpo.train(exp)
    out.update({k: np.asarray(v) for k, v in cap.items()})
    out["policy_flat_final"], out["value_flat_final"] = G.flat(pnet), G.flat(vnet)
    out["hp"] = np.asarray([damping, delta, backtracks, gae_lambda], dtype=np.float64)
    for k, val in trpo.metrics_manager.scalars.items():
        out["metric:" + k] = val
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    moved = np.abs(out["policy_flat_final"] - out["policy_flat0"]).max()
    print(name, "moved", moved, "loss trace", cap["ls_loss"][:4], "kl trace", cap["ls_kl"][:3],
          "|descent|", np.abs(cap["descent"]).max())
    assert moved == 0.0, "the case was meant to be rejected"


def sampled_actions_case(name, discrete):
    """Reference PPO.train with ONE policy step (no clip discontinuity can separate two implementations), then 1000
    single-observation draws and one batched draw under torch.manual_seed(1234)."""
    obs_dim, act_dim = (4, 3) if discrete else (17, 6)
    policy, vf, log_std = G.build(obs_dim, act_dim, discrete)
    with torch.no_grad():
        mf = (lambda o: policy.network(torch.from_numpy(o)).numpy())
        batch = (synthetic.ragged_batch(1200, obs_dim, act_dim, True, seed=8, min_len=5, max_len=90) if discrete
                 else synthetic.fixed_batch(6, 200, obs_dim, act_dim, seed=9, frac_not_done=0.34, mean_fn=mf))
    out = dict(batch)
    out["policy_flat0"], out["value_flat0"] = G.flat(policy.network), G.flat(vf.network)
    if log_std is not None:
        out["log_std"] = log_std.detach().numpy().copy()
    ppo = PPO(policy, vf, None, None, num_policy_gradients=1, num_value_gradients=1, max_kl_divergence=float("inf"))
    ppo.metrics_manager = G.Recorder()
    ppo.current_total_steps = 0
    ppo.train(Experience(**synthetic.to_experience_lists(batch, discrete)))
    out["policy_flat_final"] = G.flat(policy.network)
    rng = np.random.default_rng(21)
    probe = rng.standard_normal((1000, obs_dim)).astype(np.float32)
    torch.manual_seed(1234)
    single = np.stack([np.asarray(policy.get_action_numpy(probe[i])) for i in range(1000)])
    batched = np.asarray(policy.get_action_numpy(probe))
    out["probe_obs"], out["single_draws"], out["batched_draw"] = probe, single, batched
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, single.shape, single.dtype, batched.shape, "first", single[:5].tolist() if discrete else single[0])


def trainable_log_std_case(name):
    """PPO on a Gaussian policy whose log_std sits in the optimizer behind the network's parameters: the reference then
    trains it (autograd through Normal(loc, exp(log_std)), policies/gaussian_policy.py:25-37)."""
    set_seed_for_libraries(0)
    obs_dim, act_dim = 17, 6
    pnet, vnet = MLP([obs_dim, 64, 64, act_dim]), MLP([obs_dim, 64, 64, 1])
    log_std = torch.nn.Parameter(-0.5 * torch.ones(act_dim))
    policy = GaussianPolicy(pnet, torch.optim.Adam(list(pnet.parameters()) + [log_std], lr=3e-4), log_std)
    vf = ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3))
    with torch.no_grad():
        batch = synthetic.fixed_batch(6, 200, obs_dim, act_dim, seed=12, frac_not_done=0.34,
                                      mean_fn=lambda o: policy.network(torch.from_numpy(o)).numpy())
    out = dict(batch)
    out["policy_flat0"], out["value_flat0"] = G.flat(pnet), G.flat(vnet)
    out["log_std0"] = log_std.detach().numpy().copy()
    ppo = PPO(policy, vf, None, None, num_policy_gradients=6, num_value_gradients=3, max_kl_divergence=float("inf"))
    ppo.metrics_manager = G.Recorder()
    ppo.current_total_steps = 0
    ppo.train(Experience(**synthetic.to_experience_lists(batch, False)))
    out["policy_flat_final"], out["value_flat_final"] = G.flat(pnet), G.flat(vnet)
    out["log_std_final"] = log_std.detach().numpy().copy()
    out["old_log_std_final"] = ppo.old_policy.log_std.detach().numpy().copy()
    st = policy.optimizer.state[log_std]
    out["log_std_adam_m"], out["log_std_adam_v"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    out["log_std_adam_step"] = float(st["step"])
    for k, val in ppo.metrics_manager.scalars.items():
        out["metric:" + k] = val
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "log_std", out["log_std0"][:3], "->", out["log_std_final"][:3], "kl", out["metric:policy/kl_divergence"])


def utils_kats():
    out = {}
    rng = np.random.default_rng(3)
    x2 = rng.standard_normal((300, 3))
    out["dcs2d_in"], out["dcs2d_out"] = x2, discounted_cumulative_sums(x2, 0.9)
    x1 = rng.standard_normal(5000)
    out["dcs_long_in"], out["dcs_long_out"] = x1, discounted_cumulative_sums(x1, 0.999)
    r = rng.standard_normal(1201)
    v = rng.standard_normal(1201).astype(np.float32)
    out["gae_r"], out["gae_v"], out["gae_out"] = r, v, gae(r, 0.99, v, 0.97)
    r1 = np.asarray([2.5])  # an episode of zero steps' worth of deltas: L = 0
    out["gae_empty_out"] = gae(r1, 0.99, np.asarray([1.0], dtype=np.float32), 0.97)
    t = torch.from_numpy(rng.standard_normal(4097).astype(np.float32) * 3 + 1)
    out["norm_in"], out["norm_out"] = t.numpy().copy(), normalize_tensor(t).numpy().copy()
    set_seed_for_libraries(5)
    net = MLP([11, 64, 64, 1])
    vf = ValueFunction(net, torch.optim.Adam(net.parameters(), lr=1e-3))
    eps = [[rng.standard_normal(11).astype(np.float32) for _ in range(n)] for n in (8, 1, 33)]
    vals = compute_values(eps, vf)
    out["cv_flat_params"] = G.flat(net)
    out["cv_obs"] = np.concatenate([np.stack(e) for e in eps])
    out["cv_lengths"] = np.asarray([len(e) for e in eps])
    out["cv_out"] = np.concatenate(vals)
    boot = bootstrap_rewards_with_last_values([[1.0, 2.0], [3.0]], [True, False], [0.5, 0.25])
    out["boot_0"], out["boot_1"] = boot[0], boot[1]
    a, b = MLP([5, 7, 2]), MLP([5, 7, 2])
    out["polyak_src"], out["polyak_tgt0"] = G.flat(a), G.flat(b)
    polyak_average(a.parameters(), b.parameters(), 0.995)
    out["polyak_tgt1"] = G.flat(b)
    np.savez_compressed(os.path.join(HERE, "utils_kats.npz"), **out)
    print("utils_kats", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    torch.set_num_threads(1)
    trpo_reject_case("trpo_reject_negdamp", damping=-10.0)
    trpo_reject_case("trpo_reject_bigdelta", delta=5.0, backtracks=2)
    trpo_reject_case("trpo_reject_nan_adv", const_adv=True)
    sampled_actions_case("sampled_actions_categorical", True)
    sampled_actions_case("sampled_actions_gaussian", False)
    utils_kats()
    trainable_log_std_case("ppo_trainable_log_std")
