"""cProfile of TD3.train (config 4 shape) to see where the host time of one train() call goes."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import onpolicy as O  # noqa: E402
from rl_replicas_b200.experience import Experience  # noqa: E402
from test_gpu_offpolicy import build as build_off  # noqa: E402

rng = np.random.default_rng(2)
H = 256
mk = lambda sz: [(rng.standard_normal((o, i)).astype(np.float32) / np.sqrt(i), np.zeros(o, np.float32))
                 for i, o in zip(sz[:-1], sz[1:])]
algo, rb = build_off(True, H, O.flatten_layers(mk([11, H, H, 3])), [O.flatten_layers(mk([14, H, H, 1])),
                                                                     O.flatten_layers(mk([14, H, H, 1]))])
algo.metrics_manager = None
n_rb = 100000
ex = Experience()
obs_rb = rng.standard_normal((n_rb + 1, 11)).astype(np.float32)
ex.observations = [[obs_rb[i] for i in range(n_rb)]]
ex.actions = [[a for a in rng.uniform(-1, 1, (n_rb, 3)).astype(np.float32)]]
ex.rewards = [[float(x) for x in rng.standard_normal(n_rb)]]
ex.dones = [[bool(x) for x in (rng.random(n_rb) < 0.001)]]
ex.last_observations = [obs_rb[n_rb]]
rb.add_experience(ex)
algo.train(rb, 50, 256)
if len(sys.argv) > 1 and sys.argv[1] == "time":  # wall-clock of train() calls and of their three parts
    import time
    T = {"upload": [], "engine": [], "download": []}

    def wrap(obj, name, key):
        f = getattr(obj, name)

        def g(*a, **k):
            t0 = time.perf_counter()
            r = f(*a, **k)
            T[key].append((time.perf_counter() - t0) * 1e3)
            return r
        setattr(obj, name, g)
    wrap(algo, "_upload_state", "upload")
    wrap(algo, "_download_state", "download")
    wrap(algo._engine, "train_gather", "engine")
    for threads in (None, 1):
        if threads:
            torch.set_num_threads(threads)
        for v in T.values():
            v.clear()
        per_call = []
        for _ in range(24):
            t1 = time.perf_counter()
            algo.train(rb, 50, 256)
            per_call.append((time.perf_counter() - t1) * 1e3)
        print("torch threads", torch.get_num_threads())
        print(" per call ms:", " ".join(f"{x:.1f}" for x in per_call))
        for k, v in T.items():
            print(f" {k:9s}", " ".join(f"{x:.1f}" for x in v))
        print(f" median {np.median(per_call):.2f} ms = {50 / np.median(per_call) * 1e3:.0f} train steps/s; mean {np.mean(per_call):.2f} ms")
    sys.exit(0)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    algo.train(rb, 50, 256)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
