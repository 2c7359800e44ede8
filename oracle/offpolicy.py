"""numpy oracle for the off-policy update path (DDPG / TD3).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
Pinned against the reference's own TD3.train / DDPG.train outputs (tests/golden/td3_small.npz, ddpg_small.npz).
"ref:" paths are relative to /root/reference/src/rl_replicas/."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .onpolicy import F32, AdamState, flatten_layers, layer_sizes, mlp_backward, mlp_forward, unflatten_layers


def q_forward(q_layers, obs, act, hidden="relu"):
    """QFunction.forward: squeeze(net(cat([obs, act], -1)), -1)  (ref: q_function.py:20-32)"""
    x = np.concatenate([obs, act], axis=-1).astype(F32)
    out, acts = mlp_forward(q_layers, x, hidden, "identity")
    return out[:, 0], acts


def offpolicy_train(nets: Dict[str, list], adams: Dict[str, AdamState], minibatches: List[dict], noise, gamma=0.99,
                    rho=0.995, noise_scale=0.2, noise_clip=0.5, action_limit=1.0, policy_delay=2, twin=True,
                    p_hidden="relu", p_out="tanh", q_hidden="relu"):
    """TD3.train (ref: algorithms/td3.py:214-358) / DDPG.train (ref: algorithms/ddpg.py:195-293, twin=False, no noise,
    delay 1).  nets: policy, q1, (q2), target_policy, target_q1, (target_q2) as lists of (W, b).  Mutates nets/adams."""
    qs = ["q1", "q2"] if twin else ["q1"]
    logs = dict(q1_values=[], q2_values=[], q1_losses=[], q2_losses=[], policy_losses=[])
    for st, mb in enumerate(minibatches):
        obs, act = mb["observations"].astype(F32), mb["actions"].astype(F32)
        rew, nobs = mb["rewards"].astype(F32), mb["next_observations"].astype(F32)
        done = mb["dones"].astype(np.int32).astype(F32)
        B = obs.shape[0]
        # td3.py:325-341
        a2 = mlp_forward(nets["target_policy"], nobs, p_hidden, p_out)[0]
        if noise is not None:
            eps = np.clip(F32(noise_scale) * noise[st].astype(F32), -F32(noise_clip), F32(noise_clip))
            a2 = np.clip(a2 + eps, -F32(action_limit), F32(action_limit)).astype(F32)
        tq = q_forward(nets["target_q1"], nobs, a2, q_hidden)[0]
        if twin:
            tq = np.minimum(tq, q_forward(nets["target_q2"], nobs, a2, q_hidden)[0])
        y = (rew + F32(gamma) * (F32(1) - done) * tq).astype(F32)
        for name in qs:  # td3.py:343-358
            q, acts = q_forward(nets[name], obs, act, q_hidden)
            logs[name + "_values"].append(q.copy())
            diff = (q - y).astype(F32)
            logs[name + "_losses"].append(float(np.mean(diff.astype(np.float64) ** 2)))
            grads = mlp_backward(nets[name], acts, (F32(2) * diff / F32(B))[:, None], q_hidden, "identity")
            sizes = layer_sizes(nets[name])
            nets[name] = unflatten_layers(adams[name].apply(flatten_layers(nets[name]), flatten_layers(grads)), sizes)
        if st % policy_delay == 0:  # td3.py:244-263, 301-323
            a_pi, pacts = mlp_forward(nets["policy"], obs, p_hidden, p_out)
            q, qacts = q_forward(nets["q1"], obs, a_pi, q_hidden)
            logs["policy_losses"].append(float(-np.mean(q.astype(np.float64))))
            _, dx = mlp_backward(nets["q1"], qacts, np.full((B, 1), -1.0 / B, dtype=F32), q_hidden, "identity", need_dx=True)
            da = dx[:, obs.shape[1]:]
            pg = mlp_backward(nets["policy"], pacts, da, p_hidden, p_out)
            sizes = layer_sizes(nets["policy"])
            nets["policy"] = unflatten_layers(adams["policy"].apply(flatten_layers(nets["policy"]), flatten_layers(pg)), sizes)
            for src, dst in [("policy", "target_policy"), ("q1", "target_q1")] + ([("q2", "target_q2")] if twin else []):
                nets[dst] = [((F32(rho) * tw + F32(1.0 - rho) * w).astype(F32), (F32(rho) * tb + F32(1.0 - rho) * b).astype(F32))
                             for (w, b), (tw, tb) in zip(nets[src], nets[dst])]  # utils.py:47-57
    return logs
