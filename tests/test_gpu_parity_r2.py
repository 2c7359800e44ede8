"""Round-2 parity evidence on the GPU ("close the parity-evidence gaps"):
  (a) data-parallel update == oracle on the concatenated batch == the one-engine run (two engines on ONE GPU, a
      test-only all-reduce hook that barrier-sums their device buffers);
  (c) BASELINE config 2 at FULL size (1024 envs x 1000 steps) against the numpy oracle: scan outputs, first policy
      and value gradients, parameters after one Adam step each;
  (d) the approximate-KL trace of the 80-step policy loop against the reference's (golden), step by step;
  (e) PPO.train(Experience(nested lists)) -- the reference's actual boundary call -- against the goldens;
  (f) sampled actions: the reference's draws (golden, fixed torch seed) from the parameters a GPU train() leaves behind,
      indices bit-exact;
  (g) ConjugateGradientOptimizer corner cases by value: NaN step size -> 1.0, NaN direction -> 0, reject / restore.
"""
import threading

import numpy as np
import pytest
import torch

from conftest import batch_of, load_golden, rel_err
from oracle import onpolicy as O
from test_gpu_ppo import build, flat
from test_gpu_trpo import build_trpo

pytestmark = pytest.mark.gpu


def _layers(rng, sizes):
    return [(rng.standard_normal((o, i)).astype(np.float32) / np.sqrt(i), np.zeros(o, np.float32))
            for i, o in zip(sizes[:-1], sizes[1:])]


# ---------------------------------------------------------------------------------------------------------------
# (a) data parallel
# ---------------------------------------------------------------------------------------------------------------
class _TwoRankAllReduce:
    """Sum the buffers the two engines hand in, in rank order, and write the sum back into both -- what an
    all-reduce(sum) over two ranks does.  Both engines enqueue on the same CUDA stream, so host-side barriers are
    enough to order the device work."""

    def __init__(self):
        self.barrier = threading.Barrier(2)
        self.bufs = [None, None]
        self.calls = 0

    def hook(self, rank):
        def fn(t):
            self.bufs[rank] = t
            self.barrier.wait()
            if rank == 0:
                total = self.bufs[0] + self.bufs[1]
                self.bufs[0].copy_(total)
                self.bufs[1].copy_(total)
                self.calls += 1
            self.barrier.wait()
        return fn


@pytest.mark.parametrize("max_kl", [float("inf"), 0.002])
def test_data_parallel_update_matches_oracle_and_single_engine(max_kl):
    from rl_replicas_b200 import synthetic
    from rl_replicas_b200.engine import OLD_POLICY, POLICY, VALUE, OnPolicyEngine
    rng = np.random.default_rng(0)
    ps, vs = [17, 64, 64, 6], [17, 64, 64, 1]
    pl, vl = _layers(rng, ps), _layers(rng, vs)
    log_std = np.full(6, -0.5, np.float32)
    full = synthetic.ragged_batch(30000, 17, 6, False, seed=3, min_len=50, max_len=400,
                                  mean_fn=lambda o: O.mlp_forward(pl, o)[0])
    n = full["obs"].shape[0]
    K = 6

    def make(batch):
        e = OnPolicyEngine(ps, vs, "gaussian", batch["obs"].shape[0], batch["ep_done"].shape[0])
        e.set_params(POLICY, O.flatten_layers(pl))
        e.set_params(OLD_POLICY, O.flatten_layers(pl))
        e.set_params(VALUE, O.flatten_layers(vl))
        e.set_log_std(log_std)
        e.set_adam(POLICY, None, None, 0)
        e.set_adam(VALUE, None, None, 0)
        e.load_batch(batch)
        return e

    hp_kw = dict(max_kl_divergence=max_kl, num_policy_gradients=K, num_value_gradients=K)
    # one engine, whole batch
    single = make(full)
    st1 = single.update(OnPolicyEngine.hparams(**hp_kw))
    p1, v1 = single.get_params(POLICY), single.get_params(VALUE)
    # two engines, half the episodes each, all-reduce hook
    ar = _TwoRankAllReduce()
    engines = [make(synthetic.shard_batch(full, r, 2)) for r in range(2)]
    torch.cuda.synchronize()
    stats, errors = [None, None], []

    def work(r):
        try:
            stats[r] = engines[r].update(OnPolicyEngine.hparams(n_global_rows=n, **hp_kw), allreduce=ar.hook(r))
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
            ar.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    pd = [e.get_params(POLICY) for e in engines]
    vd = [e.get_params(VALUE) for e in engines]
    np.testing.assert_array_equal(pd[0], pd[1])  # both ranks hold bit-identical parameters
    np.testing.assert_array_equal(vd[0], vd[1])
    assert stats[0].policy_steps_applied == stats[1].policy_steps_applied == st1.policy_steps_applied
    assert ar.calls >= 1 + K  # advantage statistics + one per gradient step at least
    # the oracle on the concatenated batch
    out = O.ppo_train(full, pl, vl, "gaussian", log_std, O.AdamState(5702, 3e-4), O.AdamState(5377, 1e-3),
                      max_kl=max_kl, n_policy=K, n_value=K)
    assert stats[0].policy_steps_applied == out["policy_steps"]
    for name, got in (("dp", (pd[0], vd[0])), ("single", (p1, v1))):
        assert rel_err(got[0], out["policy_flat"]) < 1e-5, name
        assert rel_err(got[1], out["value_flat"]) < 1e-5, name
    assert rel_err(pd[0], p1) < 2e-5 and rel_err(vd[0], v1) < 2e-5  # two results, each within 1e-5 of the oracle
    assert abs(stats[0].kl_divergence - out["kl"]) < 1e-4 * abs(out["kl"]) + 1e-8
    assert abs(stats[0].adv_std - st1.adv_std) < 1e-9 * st1.adv_std
    for e in engines + [single]:
        e.close()


def test_data_parallel_peer_exchange_two_engines_one_gpu(monkeypatch):
    """The one-shot gradient exchange over peer-mapped memory (reduce_adam3 modes 3 / 4 + wait_peers_kernel), driven
    by two engines on one GPU: each on its own stream and host thread, exchange buffers attached by raw pointer (on a
    multi-GPU node the pointers come from CUDA IPC handles, tests/dist_check.py).  Same bars as the hook-based run.
    The single-launch form (mode 5) waits inside a grid that fills the device; two ranks SHARING one GPU would starve
    each other's step kernel, so this test pins the three-launch form -- mode 5 is covered by tests/dist_check.py."""
    monkeypatch.setenv("B200RL_PEER_ONE_LAUNCH", "0")
    from rl_replicas_b200 import synthetic
    from rl_replicas_b200.engine import OLD_POLICY, POLICY, VALUE, OnPolicyEngine
    rng = np.random.default_rng(1)
    ps, vs = [17, 64, 64, 6], [17, 64, 64, 1]
    pl, vl = _layers(rng, ps), _layers(rng, vs)
    log_std = np.full(6, -0.5, np.float32)
    full = synthetic.ragged_batch(20000, 17, 6, False, seed=4, min_len=50, max_len=400,
                                  mean_fn=lambda o: O.mlp_forward(pl, o)[0])
    n, K = full["obs"].shape[0], 5
    engines = []
    for r in range(2):
        b = synthetic.shard_batch(full, r, 2)
        e = OnPolicyEngine(ps, vs, "gaussian", b["obs"].shape[0], b["ep_done"].shape[0])
        e.set_params(POLICY, O.flatten_layers(pl))
        e.set_params(OLD_POLICY, O.flatten_layers(pl))
        e.set_params(VALUE, O.flatten_layers(vl))
        e.set_log_std(log_std)
        e.set_adam(POLICY, None, None, 0)
        e.set_adam(VALUE, None, None, 0)
        e.load_batch(b)
        engines.append(e)
    ptrs = [e.comm_export()[1] for e in engines]
    for r, e in enumerate(engines):
        e.comm_attach(r, ptrs)
    torch.cuda.synchronize()
    barrier, bufs, stats, errors = threading.Barrier(2), [None, None], [None, None], []

    def hook(rank):  # the three small collectives of an update; the engines run on different streams here
        def fn(t):
            torch.cuda.current_stream().synchronize()
            bufs[rank] = t
            barrier.wait()
            if rank == 0:
                total = bufs[0] + bufs[1]
                bufs[0].copy_(total)
                bufs[1].copy_(total)
                torch.cuda.current_stream().synchronize()
            barrier.wait()
        return fn

    def work(r):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                hp = OnPolicyEngine.hparams(n_global_rows=n, max_kl_divergence=float("inf"), num_policy_gradients=K,
                                            num_value_gradients=K)
                stats[r] = engines[r].update(hp, allreduce=hook(r))
                torch.cuda.current_stream().synchronize()
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
            barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert stats[0].fused == 1 and stats[1].fused == 1
    pd, vd = [e.get_params(POLICY) for e in engines], [e.get_params(VALUE) for e in engines]
    np.testing.assert_array_equal(pd[0], pd[1])
    np.testing.assert_array_equal(vd[0], vd[1])
    out = O.ppo_train(full, pl, vl, "gaussian", log_std, O.AdamState(5702, 3e-4), O.AdamState(5377, 1e-3),
                      max_kl=float("inf"), n_policy=K, n_value=K)
    assert rel_err(pd[0], out["policy_flat"]) < 1e-5 and rel_err(vd[0], out["value_flat"]) < 1e-5
    assert abs(stats[0].kl_divergence - out["kl"]) < 1e-4 * abs(out["kl"]) + 1e-8
    for e in engines:
        e.close()


# ---------------------------------------------------------------------------------------------------------------
# (c) config 2 at full size against the oracle
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_config2_against_the_oracle():
    from rl_replicas_b200 import synthetic
    E, T = 1024, 1000
    rng = np.random.default_rng(12)
    ps, vs = [17, 64, 64, 6], [17, 64, 64, 1]
    pl, vl = _layers(rng, ps), _layers(rng, vs)
    log_std = np.full(6, -0.5, np.float32)
    b = synthetic.fixed_batch(E, T, 17, 6, seed=7, frac_not_done=0.2, mean_fn=lambda o: O.mlp_forward(pl, o)[0])
    ppo = build(ps, vs, "gaussian", O.flatten_layers(pl), O.flatten_layers(vl), log_std, num_policy_gradients=1,
                num_value_gradients=1, max_kl_divergence=float("inf"))
    ppo.train_packed(b)
    # The reference sums a million per-row gradient contributions in float32 (torch's sgemm, like numpy's here): that
    # sum carries noise of its own.  Ground truth = the same float32 per-row values summed in float64; the bar for the
    # GPU is 1e-5 of max|ref| or the float32 reference arithmetic's own distance from the truth, whichever is larger.
    kw = dict(max_kl=float("inf"), n_policy=1, n_value=1)
    out = O.ppo_train(b, pl, vl, "gaussian", log_std, O.AdamState(5702, 3e-4), O.AdamState(5377, 1e-3), acc=np.float64, **kw)
    o32 = O.ppo_train(b, pl, vl, "gaussian", log_std, O.AdamState(5702, 3e-4), O.AdamState(5377, 1e-3), **kw)
    e = ppo._engine
    assert rel_err(e.view("adv_raw").cpu().numpy(), out["adv_raw"]) < 1e-5
    assert rel_err(e.view("ret").cpu().numpy(), out["ret"]) < 1e-5
    assert rel_err(e.view("values").cpu().numpy(), out["values"]) < 1e-5
    assert rel_err(e.view("old_logp").cpu().numpy(), out["old_logp"]) < 1e-5
    pg = e.view("policy_grad").cpu().numpy()[:5702]
    vg = e.view("value_grad").cpu().numpy()[:5377]
    noise_p, noise_v = rel_err(o32["grad0"], out["grad0"]), rel_err(o32["vgrad0"], out["vgrad0"])
    err_p, err_v = rel_err(pg, out["grad0"]), rel_err(vg, out["vgrad0"])
    print(f"full size: policy grad err {err_p:.2e} (float32 reference arithmetic: {noise_p:.2e}), "
          f"value grad err {err_v:.2e} ({noise_v:.2e})")
    assert err_p < max(1e-5, noise_p), (err_p, noise_p)
    assert err_v < max(1e-5, noise_v), (err_v, noise_v)
    # The first Adam step moves an entry by lr * g / (|g| + eps): entries with |g| near eps = 1e-8 turn a 1e-9 gradient
    # difference into a macroscopic fraction of lr, so at this size the parameters are checked for what they must be --
    # torch's Adam applied to the (already checked) gradient the GPU computed -- and against the oracle entry-wise
    # wherever the gradient is large enough for the step to be well conditioned.
    for got, g_gpu, flat0, n_par, lr, key, gkey in (
            (flat(ppo.policy.network), pg, O.flatten_layers(pl), 5702, 3e-4, "policy_flat", "grad0"),
            (flat(ppo.value_function.network), vg, O.flatten_layers(vl), 5377, 1e-3, "value_flat", "vgrad0")):
        want = O.AdamState(n_par, lr).apply(flat0.copy(), g_gpu)
        assert rel_err(got, want) < 1e-6, key
        solid = np.abs(out[gkey]) > 1e-4 * np.abs(out[gkey]).max()
        assert solid.mean() > 0.9
        assert rel_err(got[solid], out[key][solid]) < 1e-5, key
    st = ppo.last_update_stats
    assert abs(st.value_loss_first - out["value_loss_mean"]) < 1e-5 * out["value_loss_mean"]
    assert abs(st.kl_divergence - out["kl"]) < 1e-4 * abs(out["kl"]) + 1e-8


# ---------------------------------------------------------------------------------------------------------------
# (d) KL trace, (e) nested-list boundary call
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["ppo_gaussian_small", "ppo_categorical_cfg1", "ppo_gaussian_ragged_earlystop"])
def test_kl_trace_prefix_matches_the_reference(case):
    g = load_golden(case)
    dist = "gaussian" if "log_std" in g else "categorical"
    hp = dict(max_kl_divergence=float("inf")) if "inf" in str(g["hp_json"]) else {}
    ppo = build([int(x) for x in g["policy_sizes"]], [int(x) for x in g["value_sizes"]], dist, g["policy_flat0"],
                g["value_flat0"], g.get("log_std"), **hp)
    ppo.train_packed(batch_of(g))
    hist = ppo._engine.scalar_history()
    n = g["obs"].shape[0]
    steps = ppo.last_update_stats.policy_steps_applied
    assert steps == len(g["kl_trace"])
    # slot i + 1 carries the KL after policy step i (its forward pass runs on the updated parameters)
    kl = hist[1:steps + 1, 1] / n
    k = min(30, steps)
    ref = g["kl_trace"][:k]
    assert np.max(np.abs(kl[:k] - ref) / np.maximum(np.abs(ref), 1e-6)) < 1e-4, (kl[:k], ref)
    # value losses of the value loop, every step (smooth objective: tight)
    kv = ppo.last_update_stats.value_steps_applied
    K = ppo.num_policy_gradients
    vl = hist[K + 1:K + 1 + kv, 0] / n
    assert np.max(np.abs(vl - g["value_losses"]) / g["value_losses"]) < 1e-4


@pytest.mark.parametrize("case", ["ppo_gaussian_small", "ppo_categorical_cfg1"])
def test_train_on_nested_list_experience_matches_reference(case):
    """PPO.train(Experience(...)): the reference's boundary call (algorithms/ppo.py:139), nested Python lists in."""
    from rl_replicas_b200 import synthetic
    from rl_replicas_b200.experience import Experience
    g = load_golden(case)
    discrete = "log_std" not in g
    hp = dict(max_kl_divergence=float("inf")) if "inf" in str(g["hp_json"]) else {}
    ppo = build([int(x) for x in g["policy_sizes"]], [int(x) for x in g["value_sizes"]],
                "categorical" if discrete else "gaussian", g["policy_flat0"], g["value_flat0"], g.get("log_std"), **hp)
    ppo.train(Experience(**synthetic.to_experience_lists(batch_of(g), discrete)))
    e = ppo._engine
    assert rel_err(e.view("values").cpu().numpy(), g["values"]) < 1e-5
    assert rel_err(e.view("adv_raw").cpu().numpy(), g["adv_raw"]) < 1e-5
    assert rel_err(e.view("old_logp").cpu().numpy(), g["old_logp"]) < 1e-5
    assert ppo.last_update_stats.policy_steps_applied == len(g["kl_trace"])
    assert rel_err(flat(ppo.value_function.network), g["value_flat_final"]) < 2e-5
    assert rel_err(flat(ppo.policy.network), g["policy_flat_final"]) < 1e-2
    m = ppo.metrics_manager.s
    assert abs(m["policy/loss"] - g["metric:policy/loss"]) < 1e-6
    assert abs(m["value_function/average_loss"] - g["metric:value_function/average_loss"]) < 1e-4 * g["metric:value_function/average_loss"]


# ---------------------------------------------------------------------------------------------------------------
# (f) sampled actions
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["sampled_actions_categorical", "sampled_actions_gaussian"])
def test_sampled_actions_after_a_gpu_update_equal_the_references(case):
    g = load_golden(case)
    discrete = "log_std" not in g
    ps = [4, 64, 64, 3] if discrete else [17, 64, 64, 6]
    vs = ps[:-1] + [1]
    ppo = build(ps, vs, "categorical" if discrete else "gaussian", g["policy_flat0"], g["value_flat0"], g.get("log_std"),
                num_policy_gradients=1, num_value_gradients=1, max_kl_divergence=float("inf"))
    ppo.train_packed(batch_of(g))
    # one Adam step moves every entry by lr * g / (|g| + 1e-8): where |g| is comparable to the epsilon, a 1e-9 difference
    # in g is a visible fraction of lr -- the parameters must agree to 2 % of ONE step (6e-6 absolute), which keeps
    # the action distributions within 1e-5 of each other
    assert np.abs(flat(ppo.policy.network) - g["policy_flat_final"]).max() < 0.02 * 3e-4
    probe = g["probe_obs"]
    torch.manual_seed(1234)
    single = np.stack([np.asarray(ppo.policy.get_action_numpy(probe[i])) for i in range(probe.shape[0])])
    batched = np.asarray(ppo.policy.get_action_numpy(probe))
    if discrete:
        assert single.dtype == g["single_draws"].dtype
        np.testing.assert_array_equal(single, g["single_draws"])  # bit-exact sampled action indices
        np.testing.assert_array_equal(batched, g["batched_draw"])
    else:
        assert rel_err(single, g["single_draws"]) < 1e-5
        assert rel_err(batched, g["batched_draw"]) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# (g) conjugate-gradient optimizer corner cases, by value
# ---------------------------------------------------------------------------------------------------------------
def _reject_case(name):
    g = load_golden(name)
    damping, delta, backtracks, lam = [float(x) for x in g["hp"]]
    g["policy_sizes"], g["value_sizes"] = np.asarray([27, 64, 64, 8]), np.asarray([27, 64, 64, 1])
    trpo = build_trpo(g, num_value_gradients=2, gae_lambda=lam)
    opt = trpo.policy.optimizer
    opt.hvp_damping_coefficient, opt.max_constraint, opt.max_backtracks = damping, delta, int(backtracks)
    trpo.train_packed(batch_of(g))
    return g, trpo


def test_trpo_nan_step_size_becomes_one_and_the_ascent_step_is_rejected():
    """hvp_damping_coefficient = -10: x^T H x < 0, sqrt of a negative number, step size 1.0 by rule
    (conjugate_gradient_optimizer.py:92-93); the step then INCREASES the loss at every ratio and is rejected."""
    g, trpo = _reject_case("trpo_reject_negdamp")
    ts = trpo.last_trpo_stats
    assert ts.xhx < 0 and ts.step_size == 1.0
    assert ts.rejected == 1 and ts.accepted_index == -1
    e = trpo._engine
    assert rel_err(e.view("cg_descent").cpu().numpy(), g["descent"]) < 2e-3
    assert rel_err(e.view("cg_x").cpu().numpy(), e.view("cg_descent").cpu().numpy()) == 0.0
    np.testing.assert_array_equal(flat(trpo.policy.network), g["policy_flat0"])  # restored bit-exactly
    assert abs(ts.new_loss - g["ls_loss"][-1]) < 2e-2 * abs(g["ls_loss"][-1])  # the last ratio tried
    assert abs(ts.kl - g["ls_kl"][-1]) < 2e-2 * abs(g["ls_kl"][-1])
    assert rel_err(flat(trpo.value_function.network), g["value_flat_final"]) < 1e-5  # value steps still run


def test_trpo_trust_region_far_too_large_is_rejected():
    """delta = 5 with two backtracks: the loss rises and the KL is an order of magnitude above delta at both ratios."""
    g, trpo = _reject_case("trpo_reject_bigdelta")
    ts = trpo.last_trpo_stats
    assert ts.rejected == 1 and ts.accepted_index == -1
    assert rel_err(trpo._engine.view("cg_descent").cpu().numpy(), g["descent"]) < 2e-3
    np.testing.assert_array_equal(flat(trpo.policy.network), g["policy_flat0"])
    assert abs(ts.kl - g["ls_kl"][-1]) < 5e-2 * g["ls_kl"][-1]
    assert ts.new_loss > ts.loss_before and ts.kl > 5.0


def test_trpo_nan_direction_becomes_zero_and_nan_loss_is_rejected():
    """All advantages equal: normalize_tensor divides 0 by 0, the surrogate gradient is NaN, conjugate gradient
    returns NaN, the NaN -> 0 rule (:83) makes the descent step zero, the NaN loss rejects it (:233-237)."""
    g, trpo = _reject_case("trpo_reject_nan_adv")
    ts = trpo.last_trpo_stats
    e = trpo._engine
    assert np.all(np.isnan(g["cg_x_raw"])) and np.all(g["descent"] == 0.0)
    np.testing.assert_array_equal(e.view("cg_x").cpu().numpy(), np.zeros(g["policy_flat0"].size, np.float32))
    np.testing.assert_array_equal(e.view("cg_descent").cpu().numpy(), g["descent"])
    assert ts.rejected == 1 and np.isnan(ts.new_loss) and ts.kl == 0.0
    np.testing.assert_array_equal(flat(trpo.policy.network), g["policy_flat0"])
    assert np.isnan(trpo.metrics_manager.s["policy/loss"])
    assert rel_err(flat(trpo.value_function.network), g["value_flat_final"]) < 1e-5
