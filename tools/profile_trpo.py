"""One TRPO update at BASELINE config 3 (Ant-shaped, 1024 envs x 1000 steps) for launch lists / ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rl_replicas_b200 import synthetic  # noqa: E402

rng = np.random.default_rng(1)
ps, vs = [27, 64, 64, 8], [27, 64, 64, 1]
mk = lambda sz: [(rng.standard_normal((o, i)).astype(np.float32) / np.sqrt(i), np.zeros(o, np.float32))
                 for i, o in zip(sz[:-1], sz[1:])]
pl, vl = mk(ps), mk(vs)
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
trpo = synthetic.onpolicy_learner("trpo", pl, vl, np.full(8, -0.5, np.float32), num_value_gradients=80)
b = synthetic.fixed_batch(envs, 1000, 27, 8, seed=9, frac_not_done=0.1, mean_fn=lambda o: synthetic.numpy_mlp(pl, o))
for _ in range(2):
    trpo.train_packed(b)
torch.cuda.synchronize()
a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
trpo.train_packed(b)
z.record()
torch.cuda.synchronize()
print("TRPO update ms", a.elapsed_time(z), "launches", trpo.last_update_stats.kernel_launches, "accepted", trpo.last_trpo_stats.accepted_index)
eng = trpo._engine
hp = trpo._hparams(eng, 0)
kw = trpo.policy.optimizer.hyper_parameters()
eng.trpo_update(hp, **kw)
torch.cuda.synchronize()
a.record()
stats, ts = eng.trpo_update(hp, **kw)
z.record()
torch.cuda.synchronize()
print("TRPO update, batch resident: ms", a.elapsed_time(z), "launches", stats.kernel_launches, "fvp", ts.fvp_launches)
eng.run_stage("fvp", hp)
torch.cuda.synchronize()
a.record()
for _ in range(10):
    eng.run_stage("fvp", hp)
z.record()
torch.cuda.synchronize()
print("F v launch: ms", a.elapsed_time(z) / 10)
