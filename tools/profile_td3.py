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
if len(sys.argv) > 1 and sys.argv[1] == "time":  # wall-clock of train() calls, persistent kernel vs CUDA-graph replay
    import time
    for mega in ("1", "0"):
        os.environ["B200RL_OFFPOLICY_MEGAKERNEL"] = mega
        algo.train(rb, 50, 256)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            algo.train(rb, 50, 256)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / 20
        print(f"B200RL_OFFPOLICY_MEGAKERNEL={mega}: {ms:.3f} ms per TD3.train(rb, 50, 256) = {50 / ms * 1e3:.0f} train steps/s")
    # the engine alone: host-staged minibatches, one upload + the S steps + one read-back
    eng = algo._engine
    mbs = [rb.sample_minibatch(256) for _ in range(50)]
    st = lambda k: np.stack([np.asarray(m[k]) for m in mbs]).astype(np.float32)
    noise = torch.stack([torch.randn(256, 3) for _ in range(50)]).numpy()
    args = (algo._hparams(True, 2), st("observations"), st("actions"), st("rewards"), st("next_observations"), st("dones"), noise)
    for mega, graph in (("1", "1"), ("0", "1"), ("0", "0")):
        os.environ["B200RL_OFFPOLICY_MEGAKERNEL"], os.environ["B200RL_OFFPOLICY_GRAPH"] = mega, graph
        eng.train(*args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            out = eng.train(*args)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / 20
        print(f"engine, megakernel={mega} graph={graph}: {ms:.3f} ms per 50 steps = {50 / ms * 1e3:.0f} train steps/s; "
              f"kernel launches per call {out.get('kernel_launches')}")
    sys.exit(0)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    algo.train(rb, 50, 256)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
