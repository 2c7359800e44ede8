"""numpy oracle for the on-policy update path (PPO / VPG / TRPO preamble).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned against outputs of
the reference itself (tests/golden/*.npz, made by tests/golden/make_golden.py).

All "ref:" citations are relative to /root/reference/src/rl_replicas/.

Conventions
-----------
* An MLP is a list of (W, b) float32 pairs, W shaped [out, in] exactly like
  torch.nn.Linear (ref: networks/mlp.py:24-31).  ``flatten_layers`` gives the
  flat order W0,b0,W1,b1,... (= torch parameters_to_vector order).
* A trajectory batch is "packed": obs [N,O] f32, act [N,A] f32 (Gaussian) or
  [N] f32 (Categorical index cast to float, ref: algorithms/ppo.py:154),
  rewards [N] f64, last_obs [E,O] f32, ep_offsets [E+1] int64 (CSR), ep_done [E] bool.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

Layers = List[Tuple[np.ndarray, np.ndarray]]

F32 = np.float32
LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


# --------------------------------------------------------------------------
# MLP (ref: networks/mlp.py:24-41)
# --------------------------------------------------------------------------
def _act(z: np.ndarray, kind: str) -> np.ndarray:
    if kind == "identity":
        return z
    if kind == "tanh":
        return np.tanh(z, dtype=F32)
    if kind == "relu":
        return np.maximum(z, F32(0))
    raise ValueError(kind)


def _act_prime_from_output(a: np.ndarray, kind: str) -> np.ndarray:
    if kind == "identity":
        return np.ones_like(a)
    if kind == "tanh":
        return F32(1) - a * a
    if kind == "relu":
        return (a > 0).astype(F32)
    raise ValueError(kind)


def mlp_forward(layers: Layers, x: np.ndarray, hidden_act: str = "tanh", out_act: str = "identity"):
    """Linear -> act -> ... -> Linear -> out_act (ref: networks/mlp.py:24-41). Returns (out, activations)."""
    h = np.ascontiguousarray(x, dtype=F32)
    acts = [h]
    n = len(layers)
    for i, (w, b) in enumerate(layers):
        z = h @ w.T + b
        h = _act(z.astype(F32, copy=False), out_act if i == n - 1 else hidden_act)
        acts.append(h)
    return h, acts


def mlp_backward(layers: Layers, acts: Sequence[np.ndarray], dout: np.ndarray, hidden_act: str = "tanh",
                 out_act: str = "identity", need_dx: bool = False, acc=F32):
    """Reverse-mode gradient of sum(out * dout) w.r.t. every (W, b); restates what
    torch autograd does for the Sequential of Linear/activation (ref: ppo.py:233-235).
    ``acc=np.float64`` sums the per-row contributions of the weight / bias gradients in float64 (the per-row values
    stay the float32 ones): at a million rows a float32 sum -- torch's as much as numpy's -- carries ~1e-5 of
    summation noise of its own, and the full-size parity test needs a reference that does not."""
    n = len(layers)
    grads: Layers = [None] * n  # type: ignore
    d = dout.astype(F32, copy=False)
    for l in reversed(range(n)):
        dz = d * _act_prime_from_output(acts[l + 1], out_act if l == n - 1 else hidden_act)
        if acc is F32:
            grads[l] = ((dz.T @ acts[l]).astype(F32), dz.sum(axis=0, dtype=F32))
        else:
            grads[l] = ((dz.T.astype(acc) @ acts[l].astype(acc)).astype(F32), dz.sum(axis=0, dtype=acc).astype(F32))
        if l > 0 or need_dx:
            d = (dz @ layers[l][0]).astype(F32)
    return (grads, d) if need_dx else grads


def flatten_layers(layers: Layers) -> np.ndarray:
    return np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)]) for w, b in layers]).astype(F32)


def unflatten_layers(flat: np.ndarray, sizes: Sequence[int]) -> Layers:
    out: Layers = []
    o = 0
    for i in range(len(sizes) - 1):
        nw = sizes[i + 1] * sizes[i]
        w = flat[o:o + nw].reshape(sizes[i + 1], sizes[i]).astype(F32).copy()
        o += nw
        b = flat[o:o + sizes[i + 1]].astype(F32).copy()
        o += sizes[i + 1]
        out.append((w, b))
    assert o == flat.size
    return out


def layer_sizes(layers: Layers) -> List[int]:
    return [layers[0][0].shape[1]] + [w.shape[0] for w, _ in layers]


# --------------------------------------------------------------------------
# distributions (torch.distributions restated; call sites
# ref: policies/gaussian_policy.py:33-37, policies/categorical_policy.py:29-32)
# --------------------------------------------------------------------------
def gaussian_log_prob(mean: np.ndarray, log_std: np.ndarray, act: np.ndarray) -> np.ndarray:
    """Independent(Normal(mean, exp(log_std)), 1).log_prob(act).
    torch Normal.log_prob: -((x-mu)^2)/(2 var) - log(scale) - log(sqrt(2 pi)), summed over the last dim."""
    scale = np.exp(log_std.astype(F32))
    var = scale * scale
    log_scale = np.log(scale)
    lp = -((act - mean) ** 2) / (F32(2) * var) - log_scale - F32(LOG_SQRT_2PI)
    return lp.astype(F32).sum(axis=-1, dtype=F32)


def gaussian_entropy(log_std: np.ndarray, n_rows: int) -> np.ndarray:
    """Normal.entropy = 0.5 + 0.5 log(2 pi) + log(scale), summed over the last dim."""
    scale = np.exp(log_std.astype(F32))
    e = (F32(0.5 + 0.5 * math.log(2 * math.pi)) + np.log(scale)).astype(F32).sum(dtype=F32)
    return np.full(n_rows, e, dtype=F32)


def gaussian_kl(mean_p, log_std_p, mean_q, log_std_q) -> np.ndarray:
    """kl_divergence(Independent(Normal p), Independent(Normal q)) (ref call site: trpo.py:173)."""
    sp, sq = np.exp(log_std_p.astype(F32)), np.exp(log_std_q.astype(F32))
    var_ratio = (sp / sq) ** 2
    t1 = ((mean_p - mean_q) / sq) ** 2
    return (F32(0.5) * (var_ratio + t1 - F32(1) - np.log(var_ratio))).astype(F32).sum(axis=-1, dtype=F32)


def log_softmax(logits: np.ndarray) -> np.ndarray:
    """Categorical(logits=...) normalisation: logits - logsumexp(logits)."""
    m = logits.max(axis=-1, keepdims=True)
    z = logits - m
    return (z - np.log(np.exp(z).sum(axis=-1, keepdims=True, dtype=F32))).astype(F32)


def categorical_log_prob(logits: np.ndarray, act: np.ndarray) -> np.ndarray:
    ls = log_softmax(logits)
    idx = act.astype(np.int64).reshape(-1, 1)  # value.long() in torch
    return np.take_along_axis(ls, idx, axis=-1)[:, 0]


def categorical_entropy(logits: np.ndarray) -> np.ndarray:
    ls = log_softmax(logits)
    p = np.exp(ls)
    return -(np.maximum(ls, np.finfo(F32).min) * p).sum(axis=-1, dtype=F32)


def categorical_kl(logits_p: np.ndarray, logits_q: np.ndarray) -> np.ndarray:
    lp, lq = log_softmax(logits_p), log_softmax(logits_q)
    return (np.exp(lp) * (lp - lq)).sum(axis=-1, dtype=F32)


class Dist:
    """Tiny stand-in for the torch Distribution objects the policies return."""

    def __init__(self, kind: str, out: np.ndarray, log_std: np.ndarray | None):
        self.kind, self.out, self.log_std = kind, out, log_std

    def log_prob(self, act):
        if self.kind == "gaussian":
            return gaussian_log_prob(self.out, self.log_std, act)
        return categorical_log_prob(self.out, act)

    def entropy(self):
        if self.kind == "gaussian":
            return gaussian_entropy(self.log_std, self.out.shape[0])
        return categorical_entropy(self.out)

    def dlogp_dout(self, act) -> np.ndarray:
        """d log_prob / d network-output (what autograd yields through the distribution)."""
        if self.kind == "gaussian":
            scale = np.exp(self.log_std.astype(F32))
            return ((act - self.out) / (scale * scale)).astype(F32)
        p = np.exp(log_softmax(self.out))
        onehot = np.zeros_like(p)
        onehot[np.arange(p.shape[0]), act.astype(np.int64)] = 1
        return (onehot - p).astype(F32)


# --------------------------------------------------------------------------
# scans (ref: utils.py:14-44, 74-87; ppo.py:139-161)
# --------------------------------------------------------------------------
def discounted_cumulative_sums(x: np.ndarray, discount: float) -> np.ndarray:
    """y_t = x_t + discount * y_{t+1} in float64 (ref: utils.py:14-28; lfilter restated as a plain loop)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.empty_like(x)
    carry = 0.0
    for t in range(x.shape[0] - 1, -1, -1):
        carry = x[t] + discount * carry
        y[t] = carry
    return y


def discounted_cumulative_sums_lfilter(x: np.ndarray, discount: float) -> np.ndarray:
    """The reference's own formulation (ref: utils.py:28), used for the fast path of the CPU baseline."""
    import scipy.signal
    return np.asarray(scipy.signal.lfilter([1], [1, -discount], np.asarray(x)[::-1], axis=0)[::-1])


def gae_and_returns(rewards: np.ndarray, values: np.ndarray, last_values: np.ndarray, ep_offsets: np.ndarray,
                    ep_done: np.ndarray, gamma: float, lam: float, fast: bool = True):
    """Per episode (ref: ppo.py:142-161, utils.py:31-44, 74-87):
        R = rewards_e + [0 if done else v_L]          (float64)
        ret = dcs(R, gamma)[:-1]
        delta = R[:-1] + gamma * v[1:] - v[:-1]       (gamma*v[1:] is evaluated in float32: v is a float32 array)
        adv = dcs(delta, gamma*lam)
    ``values`` are V(obs) [N] f32, ``last_values`` V(last_obs) [E] f32.  Returns (adv_raw f32 [N], ret f32 [N])."""
    dcs = discounted_cumulative_sums_lfilter if fast else discounted_cumulative_sums
    n = int(ep_offsets[-1])
    adv = np.empty(n, dtype=np.float64)
    ret = np.empty(n, dtype=np.float64)
    for e in range(len(ep_offsets) - 1):
        s, t = int(ep_offsets[e]), int(ep_offsets[e + 1])
        v = np.concatenate([values[s:t], last_values[e:e + 1]]).astype(F32)
        boot = 0.0 if bool(ep_done[e]) else float(v[-1])
        r = np.concatenate([np.asarray(rewards[s:t], dtype=np.float64), [boot]])
        ret[s:t] = dcs(r, gamma)[:-1]
        delta = r[:-1] + (F32(gamma) * v[1:]) - v[:-1]
        adv[s:t] = dcs(delta, gamma * lam)
    return adv.astype(F32), ret.astype(F32)


def normalize(adv: np.ndarray) -> np.ndarray:
    """(x - mean) / std with the unbiased std, no epsilon (ref: utils.py:90-92).  Statistics in float64."""
    a = adv.astype(np.float64)
    mean = a.mean()
    std = math.sqrt(((a - mean) ** 2).sum() / (a.size - 1))
    return ((adv - F32(mean)) / F32(std)).astype(F32)


# --------------------------------------------------------------------------
# Adam (torch.optim.Adam single-tensor path, torch 2.5.1; call sites ref: ppo.py:235, 278)
# --------------------------------------------------------------------------
class AdamState:
    def __init__(self, n: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.step = 0
        self.m = np.zeros(n, dtype=F32)
        self.v = np.zeros(n, dtype=F32)

    def apply(self, flat: np.ndarray, grad: np.ndarray) -> np.ndarray:
        self.step += 1
        g = grad.astype(F32)
        self.m = (self.m + F32(1 - self.b1) * (g - self.m)).astype(F32)          # exp_avg.lerp_(grad, 1-beta1)
        self.v = (self.v * F32(self.b2) + F32(1 - self.b2) * (g * g)).astype(F32)  # mul_(beta2).addcmul_(g, g, 1-beta2)
        bc1 = 1 - self.b1 ** self.step
        bc2 = 1 - self.b2 ** self.step
        step_size = self.lr / bc1
        denom = (np.sqrt(self.v) / F32(math.sqrt(bc2)) + F32(self.eps)).astype(F32)
        return (flat - F32(step_size) * (self.m / denom)).astype(F32)


# --------------------------------------------------------------------------
# losses & one gradient evaluation
# --------------------------------------------------------------------------
def policy_loss_and_grad(layers: Layers, dist_kind: str, log_std, obs, act, adv, old_logp, loss_kind: str,
                         clip: float = 0.2, hidden_act: str = "tanh", n_global: int | None = None, acc=F32):
    """Returns dict(loss, grads(flat), logp, kl=mean(old_logp-logp), entropy).

    loss_kind:
      "ppo"   -mean(min(r A, clamp(r,1-c,1+c) A))            ref: ppo.py:237-257
      "vpg"   -mean(logp A)                                   ref: vpg.py:200-203
      "trpo"  -mean(r A)                                      ref: trpo.py:154-165
    The gradient w.r.t. logp is the closed form of what autograd computes."""
    n = obs.shape[0] if n_global is None else n_global
    out, acts = mlp_forward(layers, obs, hidden_act, "identity")
    d = Dist(dist_kind, out, log_std)
    logp = d.log_prob(act)
    if loss_kind == "vpg":
        loss = -np.sum(logp.astype(np.float64) * adv) / n
        coef = -adv / F32(n)
        ratio = None
    else:
        ratio = np.exp(logp - old_logp).astype(F32)
        if loss_kind == "ppo":
            s1 = ratio * adv
            s2 = np.clip(ratio, F32(1 - clip), F32(1 + clip)) * adv
            loss = -np.sum(np.minimum(s1, s2).astype(np.float64)) / n
            mask = np.where(adv >= 0, ratio <= F32(1 + clip), ratio >= F32(1 - clip)).astype(F32)
            coef = -(adv * ratio * mask) / F32(n)
        elif loss_kind == "trpo":
            loss = -np.sum((ratio * adv).astype(np.float64)) / n
            coef = -(adv * ratio) / F32(n)
        else:
            raise ValueError(loss_kind)
    dout = (coef[:, None] * d.dlogp_dout(act)).astype(F32)
    grads = mlp_backward(layers, acts, dout, hidden_act, "identity", acc=acc)
    res = {
        "loss": float(loss),
        "grad": flatten_layers(grads),
        "logp": logp,
        "entropy": d.entropy(),
    }
    if old_logp is not None:
        res["kl"] = float(np.sum((old_logp - logp).astype(np.float64)) / n)
    if dist_kind == "gaussian":
        # d logp / d log_std_a = (act_a - mu_a)^2 / var_a - 1  (torch Normal.log_prob with scale = exp(log_std),
        # ref policies/gaussian_policy.py:34): the gradient autograd hands to a log_std that sits in the optimizer
        var = np.exp(F32(2) * np.asarray(log_std, dtype=F32))
        z2 = ((act - out) ** 2 / var).astype(F32)
        res["grad_log_std"] = (coef[:, None] * (z2 - F32(1))).sum(axis=0, dtype=acc).astype(F32)
    return res


def value_loss_and_grad(layers: Layers, obs, ret, hidden_act: str = "tanh", n_global: int | None = None, acc=F32):
    """F.mse_loss(squeeze(V(obs)), ret) and its gradient (ref: ppo.py:282-287)."""
    n = obs.shape[0] if n_global is None else n_global
    out, acts = mlp_forward(layers, obs, hidden_act, "identity")
    v = out[:, 0]
    diff = (v - ret).astype(F32)
    loss = float(np.sum(diff.astype(np.float64) ** 2) / n)
    dout = (F32(2) * diff / F32(n))[:, None].astype(F32)
    grads = mlp_backward(layers, acts, dout, hidden_act, "identity", acc=acc)
    return {"loss": loss, "grad": flatten_layers(grads), "values": v}


# --------------------------------------------------------------------------
# PPO.train restated end to end (ref: algorithms/ppo.py:139-223)
# --------------------------------------------------------------------------
def ppo_train(batch: Dict[str, np.ndarray], policy: Layers, value: Layers, dist_kind: str, log_std,
              policy_adam: AdamState, value_adam: AdamState, gamma=0.99, lam=0.97, clip=0.2, max_kl=0.01,
              n_policy=80, n_value=80, old_policy: Layers | None = None, hidden_act: str = "tanh",
              trace: bool = False, acc=F32, train_log_std: bool = False) -> Dict[str, object]:
    """``train_log_std``: log_std sits behind the network's parameters in the policy optimizer (``policy_adam`` then has
    P + A entries) and is trained by the policy steps; the old policy keeps its own (initial) log_std."""
    obs, act = batch["obs"], batch["act"]
    sizes_p, sizes_v = layer_sizes(policy), layer_sizes(value)
    old_policy = policy if old_policy is None else old_policy

    # ppo.py:140-161 -- value inference, bootstrapping, returns, GAE, normalisation
    values = mlp_forward(value, obs, hidden_act)[0][:, 0]
    last_values = mlp_forward(value, batch["last_obs"], hidden_act)[0][:, 0]
    adv_raw, ret = gae_and_returns(batch["rew"], values, last_values, batch["ep_offsets"], batch["ep_done"], gamma, lam)
    adv = normalize(adv_raw)

    # ppo.py:241-243 -- old_policy is frozen during the epoch, so old log-probs are constant
    old_logp = Dist(dist_kind, mlp_forward(old_policy, obs, hidden_act)[0], log_std).log_prob(act)

    out: Dict[str, object] = {"values": values, "last_values": last_values, "adv_raw": adv_raw, "ret": ret,
                              "adv": adv, "old_logp": old_logp}
    flat_p = flatten_layers(policy)
    n_net = flat_p.size
    log_std_cur = None if log_std is None else np.asarray(log_std, dtype=F32).copy()
    if train_log_std:
        flat_p = np.concatenate([flat_p, log_std_cur])
    kl_trace, policy_params_trace = [], []
    kl = 0.0
    steps_done = 0
    for i in range(n_policy):  # ppo.py:173-181
        r = policy_loss_and_grad(unflatten_layers(flat_p[:n_net], sizes_p), dist_kind, log_std_cur, obs, act, adv, old_logp,
                                 "ppo", clip, hidden_act, acc=acc)
        if train_log_std:
            r["grad"] = np.concatenate([r["grad"], r["grad_log_std"]])
        if i == 0:  # ppo.py:164-170 (logging before the first update)
            out["loss_before"] = r["loss"]
            out["entropy_before"] = float(np.mean(r["entropy"], dtype=np.float64))
            out["logp_std_before"] = float(np.std(r["logp"].astype(np.float64), ddof=1))
            out["grad0"] = r["grad"]
        flat_p = policy_adam.apply(flat_p, r["grad"])
        if train_log_std:
            log_std_cur = flat_p[n_net:].copy()
        steps_done += 1
        if trace:
            policy_params_trace.append(flat_p.copy())
        # ppo.py:176-181 -- approx KL with the updated policy
        logp_new = Dist(dist_kind, mlp_forward(unflatten_layers(flat_p[:n_net], sizes_p), obs, hidden_act)[0],
                        log_std_cur).log_prob(act)
        kl = float(np.sum((old_logp - logp_new).astype(np.float64)) / obs.shape[0])
        kl_trace.append(kl)
        if kl > 1.5 * max_kl:
            break
    out.update(policy_flat=flat_p[:n_net], kl=kl, kl_trace=np.asarray(kl_trace), policy_steps=steps_done)
    if train_log_std:
        out["log_std"] = log_std_cur
    if trace:
        out["policy_params_trace"] = np.stack(policy_params_trace)

    flat_v = flatten_layers(value)
    vlosses = []
    for j in range(n_value):  # ppo.py:186-192
        r = value_loss_and_grad(unflatten_layers(flat_v, sizes_v), obs, ret, hidden_act, acc=acc)
        if j == 0:
            out["vgrad0"] = r["grad"]
        vlosses.append(r["loss"])
        flat_v = value_adam.apply(flat_v, r["grad"])
    out.update(value_flat=flat_v, value_losses=np.asarray(vlosses),
               value_loss_mean=float(np.mean(vlosses)) if vlosses else float("nan"))
    return out


# --------------------------------------------------------------------------
# TRPO (ref: algorithms/trpo.py:130-240, optimizers/conjugate_gradient_optimizer.py:59-250)
# --------------------------------------------------------------------------
def dist_kl(kind: str, out_old, out_new, log_std) -> np.ndarray:
    """kl_divergence(old_dist, dist) per row (ref: trpo.py:167-175)."""
    if kind == "gaussian":
        return gaussian_kl(out_old, log_std, out_new, log_std)
    return categorical_kl(out_old, out_new)


def fisher_vector_product(layers: Layers, dist_kind: str, log_std, obs, v_flat, damping: float = 1e-5,
                          hidden_act: str = "tanh") -> np.ndarray:
    """(F + damping I) v with F = (1/N) J^T M J, the Hessian of mean KL(old || new) at theta = theta_old.
    The reference obtains the same product by double backprop (conjugate_gradient_optimizer.py:133-167); at
    theta_old the two are identical.  J v by forward-mode tangents, J^T by mlp_backward."""
    sizes = layer_sizes(layers)
    vl = unflatten_layers(np.asarray(v_flat, dtype=F32), sizes)
    out, acts = mlp_forward(layers, obs, hidden_act, "identity")
    n = obs.shape[0]
    t = np.zeros((n, sizes[0]), dtype=F32)
    L = len(layers)
    for l, ((w, _), (vw, vb)) in enumerate(zip(layers, vl)):
        zt = t @ w.T + acts[l] @ vw.T + vb
        kind = "identity" if l == L - 1 else hidden_act
        t = (zt * _act_prime_from_output(acts[l + 1], kind)).astype(F32)
    if dist_kind == "gaussian":
        scale = np.exp(log_std.astype(F32))
        u = t / (scale * scale)
    else:
        p = np.exp(log_softmax(out))
        u = p * (t - (p * t).sum(axis=-1, keepdims=True))
    grads = mlp_backward(layers, acts, (u / F32(n)).astype(F32), hidden_act, "identity")
    return (flatten_layers(grads) + F32(damping) * np.asarray(v_flat, dtype=F32)).astype(F32)


def conjugate_gradient(hvp, b: np.ndarray, n_iters: int = 10, residual_tol: float = 1e-10) -> np.ndarray:
    """ref: conjugate_gradient_optimizer.py:169-202 (float32 vectors)."""
    x = np.zeros_like(b)
    r, p = b.copy(), b.copy()
    rdotr = F32(r @ r)
    for _ in range(n_iters):
        z = hvp(p)
        v = F32(rdotr / F32(p @ z))
        x = (x + v * p).astype(F32)
        r = (r - v * z).astype(F32)
        newrdotr = F32(r @ r)
        mu = F32(newrdotr / rdotr)
        p = (r + mu * p).astype(F32)
        rdotr = newrdotr
        if rdotr < residual_tol:
            break
    return x


def trpo_policy_step(layers: Layers, dist_kind: str, log_std, obs, act, adv, max_constraint=0.01, n_cg=10,
                     max_backtracks=15, backtrack_ratio=0.8, damping=1e-5, hidden_act: str = "tanh"):
    """ConjugateGradientOptimizer.step restated (ref: conjugate_gradient_optimizer.py:59-98, 204-250)."""
    sizes = layer_sizes(layers)
    out_old = mlp_forward(layers, obs, hidden_act)[0]
    old_logp = Dist(dist_kind, out_old, log_std).log_prob(act)
    g = policy_loss_and_grad(layers, dist_kind, log_std, obs, act, adv, old_logp, "trpo", hidden_act=hidden_act)
    hvp = lambda v: fisher_vector_product(layers, dist_kind, log_std, obs, v, damping, hidden_act)
    x = conjugate_gradient(hvp, g["grad"], n_cg)
    x[np.isnan(x)] = 0
    xhx = F32(x @ hvp(x))
    step_size = F32(np.sqrt(F32(2.0 * max_constraint) * (F32(1.0) / (xhx + F32(1e-8)))))
    if np.isnan(step_size):
        step_size = F32(1.0)
    descent = (step_size * x).astype(F32)
    prev = flatten_layers(layers)
    loss_before = F32(g["loss"])

    def evaluate(flat):
        lay = unflatten_layers(flat, sizes)
        o = mlp_forward(lay, obs, hidden_act)[0]
        lp = Dist(dist_kind, o, log_std).log_prob(act)
        loss = F32(-np.mean((np.exp(lp - old_logp) * adv).astype(np.float64)))
        kl = F32(np.mean(dist_kl(dist_kind, out_old, o, log_std).astype(np.float64)))
        return loss, kl

    accepted, new_loss, kl, flat = -1, None, None, prev
    for k in range(max_backtracks):
        flat = (prev - F32(backtrack_ratio ** k) * descent).astype(F32)
        new_loss, kl = evaluate(flat)
        if new_loss < loss_before and kl <= max_constraint:
            accepted = k
            break
    rejected = bool(np.isnan(new_loss) or np.isnan(kl) or new_loss >= loss_before or kl >= max_constraint)
    if rejected:
        flat = prev
    return dict(grad=g["grad"], x=x, step_size=float(step_size), xhx=float(xhx), descent=descent, policy_flat=flat,
                accepted=accepted, rejected=rejected, loss_before=float(loss_before), new_loss=float(new_loss),
                kl=float(kl), entropy=float(np.mean(g["entropy"], dtype=np.float64)),
                logp_std=float(np.std(g["logp"].astype(np.float64), ddof=1)))
