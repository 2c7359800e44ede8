"""Prints the parity margins of the PPO path (goldens + oracle cases) for the library in use (B200RL_LIB selects an A/B
build): policy / value errors after full updates, KL-trace and value-loss-trace errors, first gradients."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import batch_of, load_golden, rel_err  # noqa: E402
from oracle import onpolicy as O  # noqa: E402
from test_gpu_ppo import build, flat  # noqa: E402
from rl_replicas_b200 import synthetic  # noqa: E402

for case in ("ppo_gaussian_small", "ppo_categorical_cfg1", "ppo_gaussian_ragged_earlystop"):
    g = load_golden(case)
    dist = "gaussian" if "log_std" in g else "categorical"
    hp = dict(max_kl_divergence=float("inf")) if "inf" in str(g["hp_json"]) else {}
    ppo = build([int(x) for x in g["policy_sizes"]], [int(x) for x in g["value_sizes"]], dist, g["policy_flat0"],
                g["value_flat0"], g.get("log_std"), **hp)
    ppo.train_packed(batch_of(g))
    hist, n = ppo._engine.scalar_history(), g["obs"].shape[0]
    steps = ppo.last_update_stats.policy_steps_applied
    kl = hist[1:steps + 1, 1] / n
    k = min(30, steps)
    K = ppo.num_policy_gradients
    vl = hist[K + 1:K + 1 + 80, 0] / n
    print(f"{case}: fused={ppo.last_update_stats.fused} steps {steps}/{len(g['kl_trace'])} "
          f"policy {rel_err(flat(ppo.policy.network), g['policy_flat_final']):.2e} "
          f"value {rel_err(flat(ppo.value_function.network), g['value_flat_final']):.2e} "
          f"kl-trace {np.max(np.abs(kl[:k] - g['kl_trace'][:k]) / np.maximum(np.abs(g['kl_trace'][:k]), 1e-6)):.2e} "
          f"vloss-trace {np.max(np.abs(vl - g['value_losses']) / g['value_losses']):.2e}")
    one = build([int(x) for x in g["policy_sizes"]], [int(x) for x in g["value_sizes"]], dist, g["policy_flat0"],
                g["value_flat0"], g.get("log_std"), num_policy_gradients=1, num_value_gradients=1,
                max_kl_divergence=float("inf"))
    one.train_packed(batch_of(g))
    e = one._engine
    print(f"   one step: grad0 {rel_err(e.view('policy_grad').cpu().numpy()[:g['grad0'].size], g['grad0']):.2e} "
          f"vgrad0 {rel_err(e.view('value_grad').cpu().numpy()[:g['vgrad0'].size], g['vgrad0']):.2e} "
          f"policy1 {rel_err(flat(one.policy.network), g['policy_flat1']):.2e} value1 {rel_err(flat(one.value_function.network), g['value_flat1']):.2e} "
          f"values {rel_err(e.view('values').cpu().numpy(), g['values']):.2e} old_logp {rel_err(e.view('old_logp').cpu().numpy(), g['old_logp']):.2e}")

rng = np.random.default_rng(3)
ps, vs = [17, 64, 64, 6], [17, 64, 64, 1]
mk = lambda sz: [(rng.standard_normal((o, i)).astype(np.float32) / np.sqrt(i), 0.1 * rng.standard_normal(o).astype(np.float32))
                 for i, o in zip(sz[:-1], sz[1:])]
pl, vl = mk(ps), mk(vs)
log_std = np.full(6, -0.5, np.float32)
for envs, T, K in ((64, 250, 5), (256, 500, 20)):
    b = synthetic.fixed_batch(envs, T, 17, 6, seed=5, frac_not_done=0.1, mean_fn=lambda o: O.mlp_forward(pl, o)[0])
    ppo = build(ps, vs, "gaussian", O.flatten_layers(pl), O.flatten_layers(vl), log_std, num_policy_gradients=K,
                num_value_gradients=K, max_kl_divergence=float("inf"))
    ppo.train_packed(b)
    out = O.ppo_train(b, pl, vl, "gaussian", log_std, O.AdamState(5702, 3e-4), O.AdamState(5377, 1e-3),
                      max_kl=float("inf"), n_policy=K, n_value=K)
    print(f"oracle {envs}x{T}, {K}+{K} steps: policy {rel_err(flat(ppo.policy.network), out['policy_flat']):.2e} "
          f"value {rel_err(flat(ppo.value_function.network), out['value_flat']):.2e} "
          f"kl {abs(ppo.last_update_stats.kl_divergence - out['kl']) / abs(out['kl']):.2e}")
