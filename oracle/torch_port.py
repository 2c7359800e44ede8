"""torch-CPU port of the reference's PPO.train for the CPU-baseline / `--impl reference` arm.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- see oracle/__init__.py.  The reference (rl_replicas) is Python on torch CPU
ops and cannot travel to the GPU box; this port issues the SAME torch calls on the same packed inputs
(F.linear / tanh, torch.distributions log_prob, autograd backward, torch.optim.Adam, scipy lfilter scans), so its
speed on the host cores is representative of the reference's own update path (ref: algorithms/ppo.py:139-287,
utils.py:14-92) and its results track the reference to float32 round-off (pinned in tests/test_oracle_golden.py).
The reference's Python-list packing overhead (compute_values' np.stack, ~18 % of its time) is NOT included: the
port starts from packed arrays, which favours the CPU arm.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import scipy.signal
import torch
from torch.distributions import Categorical, Independent, Normal


def _mlp(layers):
    mods = []
    for i, (w, b) in enumerate(layers):
        lin = torch.nn.Linear(w.shape[1], w.shape[0])
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(w))
            lin.bias.copy_(torch.from_numpy(b))
        mods += [lin, torch.nn.Tanh() if i < len(layers) - 1 else torch.nn.Identity()]
    return torch.nn.Sequential(*mods)


def _dcs(x, c):  # ref utils.py:28
    return scipy.signal.lfilter([1], [1, -c], x[::-1], axis=0)[::-1]


def _flat(net):
    return torch.nn.utils.parameters_to_vector(net.parameters()).detach().numpy().copy()


def ppo_train(batch: Dict[str, np.ndarray], policy_layers, value_layers, dist_kind: str, log_std, gamma=0.99, lam=0.97,
              clip=0.2, max_kl=0.01, n_policy=80, n_value=80, policy_lr=3e-4, value_lr=1e-3):
    pnet, vnet = _mlp(policy_layers), _mlp(value_layers)
    old = _mlp(policy_layers)
    popt = torch.optim.Adam(pnet.parameters(), lr=policy_lr)
    vopt = torch.optim.Adam(vnet.parameters(), lr=value_lr)
    ls = None if log_std is None else torch.from_numpy(np.asarray(log_std, dtype=np.float32))

    def dist(net, o):
        out = net(o)
        return Independent(Normal(out, torch.exp(ls)), 1) if dist_kind == "gaussian" else Categorical(logits=out)

    obs, act = torch.from_numpy(batch["obs"]), torch.from_numpy(batch["act"])
    off, done = batch["ep_offsets"], batch["ep_done"]
    with torch.no_grad():  # ppo.py:140-161
        v = vnet(obs).flatten().numpy()
        vl = vnet(torch.from_numpy(batch["last_obs"])).flatten().numpy()
    rets, advs = [], []
    for e in range(len(done)):
        s, t = int(off[e]), int(off[e + 1])
        ve = np.concatenate([v[s:t], vl[e:e + 1]])
        r = np.concatenate([batch["rew"][s:t], [0.0 if done[e] else float(ve[-1])]])
        rets.append(_dcs(r, gamma)[:-1])
        advs.append(_dcs(r[:-1] + gamma * ve[1:] - ve[:-1], gamma * lam))
    ret = torch.from_numpy(np.concatenate(rets)).float()
    adv = torch.from_numpy(np.concatenate(advs)).float()
    adv = (adv - adv.mean()) / adv.std()
    with torch.no_grad():
        old_lp = dist(old, obs).log_prob(act)
    kl = torch.zeros(())
    steps = 0
    for _ in range(n_policy):  # ppo.py:173-181
        lp = dist(pnet, obs).log_prob(act)
        ratio = torch.exp(lp - old_lp)
        loss = -torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv).mean()
        popt.zero_grad()
        loss.backward()
        popt.step()
        steps += 1
        with torch.no_grad():
            kl = (old_lp - dist(pnet, obs).log_prob(act)).mean()
        if kl > 1.5 * max_kl:
            break
    vlosses = []
    for _ in range(n_value):  # ppo.py:186-192
        vloss = torch.nn.functional.mse_loss(vnet(obs).squeeze(-1), ret)
        vopt.zero_grad()
        vloss.backward()
        vopt.step()
        vlosses.append(float(vloss.detach()))
    return dict(policy_flat=_flat(pnet), value_flat=_flat(vnet), kl=float(kl), policy_steps=steps,
                value_losses=np.asarray(vlosses), ret=ret.numpy(), adv=adv.numpy())
