"""Generate golden vectors by running the UNMODIFIED reference (rl_replicas @ /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
gymnasium is not installed; it is used on the update path for type annotations only
(SURVEY.md section 8c), so a stub module is injected before the import.

Each .npz holds the packed inputs, the initial flat parameters and what the reference
computed: values, returns, raw/normalised advantages, log-probs, first-step gradients,
parameters after 1 step and after the full PPO.train, Adam state, KL trace, value losses
and the five logged scalars.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _stub_gymnasium():
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Env:  # noqa: D401
        pass

    class Space:
        pass

    class Box(Space):
        pass

    class Discrete(Space):
        pass

    gym.Env, gym.Space, gym.spaces, gym.make = Env, Space, spaces, (lambda *a, **k: None)
    spaces.Box, spaces.Discrete = Box, Discrete
    sys.modules["gymnasium"], sys.modules["gymnasium.spaces"] = gym, spaces


_stub_gymnasium()
sys.path.insert(0, "/root/reference/src")
import rl_replicas.algorithms.ppo as ref_ppo_mod  # noqa: E402
import rl_replicas.algorithms.vpg as ref_vpg_mod  # noqa: E402
from rl_replicas.algorithms import PPO, TRPO, VPG  # noqa: E402
from rl_replicas.optimizers import ConjugateGradientOptimizer  # noqa: E402
from rl_replicas.experience import Experience  # noqa: E402
from rl_replicas.networks import MLP  # noqa: E402
from rl_replicas.policies import CategoricalPolicy, GaussianPolicy  # noqa: E402
from rl_replicas.utils import discounted_cumulative_sums, gae, set_seed_for_libraries  # noqa: E402
from rl_replicas.value_function import ValueFunction  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "b200_synthetic", os.path.join(ROOT, "reinforcement-learning-replications_b200", "synthetic.py"))
synthetic = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synthetic)


class Recorder:
    def __init__(self):
        self.scalars = {}

    def record_scalar(self, tag, scalar, total_steps=None, tensorboard=False):
        self.scalars[tag] = float(scalar)


def flat(module):
    return torch.nn.utils.parameters_to_vector(module.parameters()).detach().numpy().copy()


def flat_grad(module):
    return torch.cat([p.grad.reshape(-1) for p in module.parameters()]).detach().numpy().copy()


def adam_state(opt):
    ps = [p for g in opt.param_groups for p in g["params"]]
    m = np.concatenate([opt.state[p]["exp_avg"].reshape(-1).numpy() for p in ps])
    v = np.concatenate([opt.state[p]["exp_avg_sq"].reshape(-1).numpy() for p in ps])
    step = float(opt.state[ps[0]]["step"])
    return m, v, step


def build(obs_dim, act_dim, discrete, hidden=(64, 64)):
    set_seed_for_libraries(0)
    pnet = MLP([obs_dim, *hidden, act_dim])
    vnet = MLP([obs_dim, *hidden, 1])
    if discrete:
        policy = CategoricalPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=3e-4))
        log_std = None
    else:
        log_std = torch.nn.Parameter(-0.5 * torch.ones(act_dim))
        policy = GaussianPolicy(pnet, torch.optim.Adam(pnet.parameters(), lr=3e-4), log_std)
    vf = ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3))
    return policy, vf, log_std


def run_ppo_case(name, batch_fn, obs_dim, act_dim, discrete, **hp):
    policy, vf, log_std = build(obs_dim, act_dim, discrete)
    with torch.no_grad():
        mean_fn = (lambda o: policy.network(torch.from_numpy(o)).numpy())
        batch = batch_fn(mean_fn)
    exp = Experience(**synthetic.to_experience_lists(batch, discrete))
    ppo = PPO(policy, vf, None, None, **hp)
    ppo.metrics_manager = Recorder()
    ppo.current_total_steps = 0

    out = dict(batch)
    out["policy_flat0"] = flat(policy.network)
    out["value_flat0"] = flat(vf.network)
    if log_std is not None:
        out["log_std"] = log_std.detach().numpy().copy()
    out["policy_sizes"] = np.asarray([obs_dim, 64, 64, act_dim])
    out["value_sizes"] = np.asarray([obs_dim, 64, 64, 1])

    # --- capture intermediates by wrapping names bound inside the reference's ppo module
    cap = {}
    orig_norm = ref_ppo_mod.normalize_tensor
    orig_gae = ref_ppo_mod.gae
    orig_dcs = ref_ppo_mod.discounted_cumulative_sums
    orig_cv = ref_ppo_mod.compute_values
    gaes, rets = [], []

    def norm_wrap(t):
        cap["adv_raw"] = t.numpy().copy()
        r = orig_norm(t)
        cap["adv"] = r.numpy().copy()
        return r

    def gae_wrap(*a):
        r = orig_gae(*a)
        gaes.append(r)
        return r

    def dcs_wrap(v, d):
        r = orig_dcs(v, d)
        rets.append(r[:-1])
        return r

    def cv_wrap(o, v):
        r = orig_cv(o, v)
        cap["values_with_last"] = [x.copy() for x in r]
        return r

    ref_ppo_mod.normalize_tensor, ref_ppo_mod.gae = norm_wrap, gae_wrap
    ref_ppo_mod.discounted_cumulative_sums, ref_ppo_mod.compute_values = dcs_wrap, cv_wrap

    kl_trace, ptrace = [], []
    orig_tp, orig_kl, orig_tv = ppo.train_policy, ppo.compute_approximate_kl_divergence, ppo.train_value_function

    def tp_wrap(o, a, adv):
        if "old_logp" not in cap:
            with torch.no_grad():
                cap["old_logp"] = ppo.old_policy(o).log_prob(a).numpy().copy()
        orig_tp(o, a, adv)
        if "grad0" not in cap:
            cap["grad0"] = flat_grad(policy.network)
            cap["policy_flat1"] = flat(policy.network)

    def kl_wrap(o, a):
        r = orig_kl(o, a)
        kl_trace.append(float(r))
        return r

    vlosses = []

    def tv_wrap(o, r):
        loss = orig_tv(o, r)
        if "vgrad0" not in cap:
            cap["vgrad0"] = flat_grad(vf.network)
            cap["value_flat1"] = flat(vf.network)
        vlosses.append(float(loss))
        return loss

    ppo.train_policy, ppo.compute_approximate_kl_divergence, ppo.train_value_function = tp_wrap, kl_wrap, tv_wrap
    try:
        ppo.train(exp)
    finally:
        ref_ppo_mod.normalize_tensor, ref_ppo_mod.gae = orig_norm, orig_gae
        ref_ppo_mod.discounted_cumulative_sums, ref_ppo_mod.compute_values = orig_dcs, orig_cv

    vwl = cap.pop("values_with_last")
    out["values"] = np.concatenate([v[:-1] for v in vwl]).astype(np.float32)
    out["last_values"] = np.asarray([v[-1] for v in vwl], dtype=np.float32)
    # dcs is called once per episode for returns (ppo.py:148-151) and once inside every gae() call
    e = len(batch["ep_done"])
    out["ret"] = np.concatenate(rets[:e]).astype(np.float32)
    out["adv_raw64"] = np.concatenate(gaes)
    out.update(cap)
    out["kl_trace"] = np.asarray(kl_trace)
    out["value_losses"] = np.asarray(vlosses)
    out["policy_flat_final"] = flat(policy.network)
    out["value_flat_final"] = flat(vf.network)
    out["old_policy_flat_final"] = flat(ppo.old_policy.network)
    m, v, step = adam_state(policy.optimizer)
    out["policy_adam_m"], out["policy_adam_v"], out["policy_adam_step"] = m, v, step
    m, v, step = adam_state(vf.optimizer)
    out["value_adam_m"], out["value_adam_v"], out["value_adam_step"] = m, v, step
    for k, val in ppo.metrics_manager.scalars.items():
        out["metric:" + k] = val
    out["hp_json"] = np.asarray(repr(sorted(hp.items())))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "N=", batch["obs"].shape[0], "E=", e, "policy steps=", len(kl_trace), "kl=", kl_trace[-1],
          "->", os.path.getsize(path) // 1024, "KiB")


def run_vpg_case(name, batch_fn, obs_dim, act_dim, discrete):
    policy, vf, log_std = build(obs_dim, act_dim, discrete)
    with torch.no_grad():
        batch = batch_fn(lambda o: policy.network(torch.from_numpy(o)).numpy())
    exp = Experience(**synthetic.to_experience_lists(batch, discrete))
    vpg = VPG(policy, vf, None, None, num_value_gradients=5)
    vpg.metrics_manager = Recorder()
    vpg.current_total_steps = 0
    out = dict(batch)
    out["policy_flat0"], out["value_flat0"] = flat(policy.network), flat(vf.network)
    if log_std is not None:
        out["log_std"] = log_std.detach().numpy().copy()
    out["policy_sizes"] = np.asarray([obs_dim, 64, 64, act_dim])
    out["value_sizes"] = np.asarray([obs_dim, 64, 64, 1])
    vpg.train(exp)
    out["grad0"] = flat_grad(policy.network)
    out["policy_flat_final"], out["value_flat_final"] = flat(policy.network), flat(vf.network)
    for k, val in vpg.metrics_manager.scalars.items():
        out["metric:" + k] = val
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "done")


def run_trpo_case(name, batch_fn, obs_dim, act_dim, discrete):
    """TRPO.train with the reference's ConjugateGradientOptimizer; records g, one Hessian-vector product of a fixed
    probe vector, the CG solution, the descent step and the outcome of the line search."""
    set_seed_for_libraries(0)
    pnet, vnet = MLP([obs_dim, 64, 64, act_dim]), MLP([obs_dim, 64, 64, 1])
    opt = ConjugateGradientOptimizer(pnet.parameters())
    if discrete:
        policy, log_std = CategoricalPolicy(pnet, opt), None
    else:
        log_std = torch.nn.Parameter(-0.5 * torch.ones(act_dim))
        policy = GaussianPolicy(pnet, opt, log_std)
    vf = ValueFunction(vnet, torch.optim.Adam(vnet.parameters(), lr=1e-3))
    with torch.no_grad():
        batch = batch_fn(lambda o: policy.network(torch.from_numpy(o)).numpy())
    exp = Experience(**synthetic.to_experience_lists(batch, discrete))
    trpo = TRPO(policy, vf, None, None, num_value_gradients=5)
    trpo.metrics_manager = Recorder()
    trpo.current_total_steps = 0
    out = dict(batch)
    out["policy_flat0"], out["value_flat0"] = flat(policy.network), flat(vf.network)
    if log_std is not None:
        out["log_std"] = log_std.detach().numpy().copy()
    out["policy_sizes"] = np.asarray([obs_dim, 64, 64, act_dim])
    out["value_sizes"] = np.asarray([obs_dim, 64, 64, 1])
    cap = {}
    rng = np.random.default_rng(11)
    probe = rng.standard_normal(out["policy_flat0"].size).astype(np.float32)
    out["hvp_probe"] = probe
    orig_build, orig_cg, orig_ls = opt._build_hessian_vector_product, opt._conjugate_gradient, opt._backtracking_line_search

    def build_wrap(fn, params):
        hvp = orig_build(fn, params)
        if "hvp_of_probe" not in cap:
            cap["hvp_of_probe"] = hvp(torch.from_numpy(probe)).detach().numpy().copy()
        return hvp

    def cg_wrap(hvp, b, residual_tol=1e-10):
        cap["grad0"] = b.detach().numpy().copy()
        x = orig_cg(hvp, b, residual_tol)
        cap["cg_x"] = x.detach().numpy().copy()
        return x

    def ls_wrap(params, descent_step, loss_fn, kl_fn):
        cap["descent"] = torch.as_tensor(descent_step).detach().numpy().copy()
        orig_ls(params, descent_step, loss_fn, kl_fn)
        with torch.no_grad():
            cap["final_loss"] = float(loss_fn())
            cap["final_kl"] = float(kl_fn())

    opt._build_hessian_vector_product, opt._conjugate_gradient, opt._backtracking_line_search = build_wrap, cg_wrap, ls_wrap
    import rl_replicas.algorithms.trpo as ref_trpo_mod
    orig_norm = ref_trpo_mod.normalize_tensor

    def norm_wrap(t):
        r = orig_norm(t)
        cap["adv"] = r.numpy().copy()
        return r

    ref_trpo_mod.normalize_tensor = norm_wrap
    try:
        trpo.train(exp)
    finally:
        ref_trpo_mod.normalize_tensor = orig_norm
    out.update(cap)
    out["policy_flat_final"], out["value_flat_final"] = flat(policy.network), flat(vf.network)
    for k, val in trpo.metrics_manager.scalars.items():
        out["metric:" + k] = val
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    moved = np.abs(out["policy_flat_final"] - out["policy_flat0"]).max()
    print(name, "N=", batch["obs"].shape[0], "moved", moved, "final kl", cap["final_kl"], "loss", cap["final_loss"])


def scan_kats():
    """Known-answer tests of the two scan helpers, taken from the reference functions themselves."""
    out = {}
    out["dcs_in"] = np.asarray([1.0, 2.0, 3.0])
    out["dcs_out"] = discounted_cumulative_sums(out["dcs_in"], 0.5)
    r = np.asarray([1.0, 1.0, 1.0])
    v = np.asarray([0.5, 0.4, 0.3, 7.0], dtype=np.float32)
    for tag, boot in (("done", 0.0), ("notdone", float(v[-1]))):
        rb = np.concatenate([r, [boot]])
        out["gae_" + tag] = gae(rb, 0.99, v, 0.97)
        out["ret_" + tag] = discounted_cumulative_sums(rb, 0.99)[:-1]
    out["kat_r"], out["kat_v"] = r, v
    rng = np.random.default_rng(7)
    x = rng.standard_normal(777)
    out["dcs_rand_in"], out["dcs_rand_out"] = x, discounted_cumulative_sums(x, 0.99 * 0.97)
    np.savez_compressed(os.path.join(HERE, "scan_kats.npz"), **out)
    print("scan_kats", out["dcs_out"], out["gae_done"], out["ret_done"], out["ret_notdone"])


if __name__ == "__main__":
    torch.set_num_threads(1)  # deterministic summation order for the fixtures
    scan_kats()
    # config 1 stand-in: CartPole-shaped, Categorical, ragged, reward 1.0, defaults (early stop on)
    run_ppo_case("ppo_categorical_cfg1",
                 lambda mf: synthetic.ragged_batch(4000, 4, 2, True, seed=0, reward_const=1.0), 4, 2, True)
    # config 2 stand-in, small: HalfCheetah-shaped Gaussian, 6 envs x 200, 1/3 of episodes not done; fixed 80+80 steps
    run_ppo_case("ppo_gaussian_small",
                 lambda mf: synthetic.fixed_batch(6, 200, 17, 6, seed=0, frac_not_done=0.34, mean_fn=mf), 17, 6, False,
                 max_kl_divergence=float("inf"))
    # same shape, default KL early stop, ragged episodes
    run_ppo_case("ppo_gaussian_ragged_earlystop",
                 lambda mf: synthetic.ragged_batch(1500, 17, 6, False, seed=1, min_len=1, max_len=120, mean_fn=mf),
                 17, 6, False)
    run_trpo_case("trpo_gaussian_small",
                  lambda mf: synthetic.fixed_batch(8, 150, 27, 8, seed=4, frac_not_done=0.3, mean_fn=mf), 27, 8, False)
    run_trpo_case("trpo_categorical_small",
                  lambda mf: synthetic.ragged_batch(1000, 4, 3, True, seed=5, min_len=5, max_len=80), 4, 3, True)
    run_vpg_case("vpg_gaussian_small",
                 lambda mf: synthetic.fixed_batch(4, 150, 17, 6, seed=2, frac_not_done=0.5, mean_fn=mf), 17, 6, False)
