"""CPU test of the data-parallel decomposition (world_size 2, gloo): shard by environment, local sums scaled by the
GLOBAL row count, all-reduce(sum) of [gradient + scalar tail] and of the 3 advantage statistics == single-process
result; likewise TRPO's surrogate gradient, Fisher-vector product and line-search KL (b200rl_trpo_update_dp).  This is
the host-side contract the engine's all-reduce callback implements; the
arithmetic inside each rank is the numpy oracle here (the CUDA path is covered by tests/dist_check.py on GPUs)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from oracle import onpolicy as O
    from rl_replicas_b200 import synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    ps, vs = [17, 64, 64, 6], [17, 64, 64, 1]
    mk = lambda sz: [(rng.standard_normal((o, i)).astype(np.float32) / np.sqrt(i), np.zeros(o, np.float32))
                     for i, o in zip(sz[:-1], sz[1:])]
    pl, vl = mk(ps), mk(vs)
    log_std = np.full(6, -0.5, np.float32)
    full = synthetic.ragged_batch(3000, 17, 6, False, seed=3, min_len=5, max_len=90, mean_fn=lambda o: O.mlp_forward(pl, o)[0])
    b = synthetic.shard_batch(full, rank, world)
    assert int(b["ep_offsets"][-1]) == b["obs"].shape[0]

    def local_preamble(batch):
        v = O.mlp_forward(vl, batch["obs"])[0][:, 0]
        lv = O.mlp_forward(vl, batch["last_obs"])[0][:, 0]
        return O.gae_and_returns(batch["rew"], v, lv, batch["ep_offsets"], batch["ep_done"], 0.99, 0.97)

    adv_raw, ret = local_preamble(b)
    a64 = adv_raw.astype(np.float64)
    stats = torch.tensor([a64.sum(), (a64 ** 2).sum(), float(a64.size)], dtype=torch.float64)
    dist.all_reduce(stats)  # (2) one 3-scalar all-reduce per update
    n_glob = int(stats[2])
    mean = float(stats[0] / stats[2])
    std = float(np.sqrt((stats[1] - stats[2] * mean * mean) / (stats[2] - 1)))
    adv = ((adv_raw - np.float32(mean)) / np.float32(std)).astype(np.float32)
    old_logp = O.Dist("gaussian", O.mlp_forward(pl, b["obs"])[0], log_std).log_prob(b["act"])
    r = O.policy_loss_and_grad(pl, "gaussian", log_std, b["obs"], b["act"], adv, old_logp, "ppo", 0.2, n_global=n_glob)
    buf = torch.from_numpy(np.concatenate([r["grad"], np.asarray([r["loss"] * n_glob, r["kl"] * n_glob], np.float32)]))
    dist.all_reduce(buf)  # (1) one all-reduce per gradient step: gradient + piggy-backed scalar sums
    rv = O.value_loss_and_grad(vl, b["obs"], ret, n_global=n_glob)
    vbuf = torch.from_numpy(rv["grad"].copy())
    dist.all_reduce(vbuf)
    # TRPO (b200rl_trpo_update_dp): the surrogate gradient, a Fisher-vector product and the true KL of a line-search
    # evaluation are sums over rows too -- local sums scaled by the GLOBAL row count, one all-reduce each
    rt = O.policy_loss_and_grad(pl, "gaussian", log_std, b["obs"], b["act"], adv, old_logp, "trpo", 0.2, n_global=n_glob)
    tbuf = torch.from_numpy(np.concatenate([rt["grad"], np.asarray([rt["loss"] * n_glob], np.float32)]))
    dist.all_reduce(tbuf)
    probe = np.random.default_rng(5).standard_normal(rt["grad"].size).astype(np.float32)
    n_loc = b["obs"].shape[0]
    # the oracle's F v divides by its own row count and adds damping * v: undo both, rescale by 1 / N_global
    fv_local = (O.fisher_vector_product(pl, "gaussian", log_std, b["obs"], probe, damping=0.0) * np.float32(n_loc / n_glob))
    fbuf = torch.from_numpy(fv_local.astype(np.float32).copy())
    dist.all_reduce(fbuf)
    moved = [(w + 0.01 * np.random.default_rng(6).standard_normal(w.shape).astype(np.float32), bb) for w, bb in pl]
    kl_local = O.dist_kl("gaussian", O.mlp_forward(pl, b["obs"])[0], O.mlp_forward(moved, b["obs"])[0], log_std)
    kbuf = torch.tensor([float(kl_local.astype(np.float64).sum())], dtype=torch.float64)
    dist.all_reduce(kbuf)
    if rank == 0:
        rtf_adv, _ = local_preamble(full)
        rtf_old = O.Dist("gaussian", O.mlp_forward(pl, full["obs"])[0], log_std).log_prob(full["act"])
        rtf = O.policy_loss_and_grad(pl, "gaussian", log_std, full["obs"], full["act"], O.normalize(rtf_adv), rtf_old,
                                     "trpo", 0.2)
        tg = tbuf.numpy()
        out["trpo_grad_err"] = float(np.abs(tg[:-1] - rtf["grad"]).max() / np.abs(rtf["grad"]).max())
        out["trpo_loss_err"] = float(abs(tg[-1] / n_glob - rtf["loss"]))
        fv_full = O.fisher_vector_product(pl, "gaussian", log_std, full["obs"], probe, damping=0.0)
        out["fvp_err"] = float(np.abs(fbuf.numpy() - fv_full).max() / np.abs(fv_full).max())
        kl_full = O.dist_kl("gaussian", O.mlp_forward(pl, full["obs"])[0], O.mlp_forward(moved, full["obs"])[0], log_std)
        out["kl_err"] = float(abs(kbuf.item() / n_glob - kl_full.astype(np.float64).mean()) / kl_full.mean())
        adv_f, ret_f = local_preamble(full)
        adv_n = O.normalize(adv_f)
        old_f = O.Dist("gaussian", O.mlp_forward(pl, full["obs"])[0], log_std).log_prob(full["act"])
        rf = O.policy_loss_and_grad(pl, "gaussian", log_std, full["obs"], full["act"], adv_n, old_f, "ppo", 0.2)
        rvf = O.value_loss_and_grad(vl, full["obs"], ret_f)
        g = buf.numpy()
        out["grad_err"] = float(np.abs(g[:-2] - rf["grad"]).max() / np.abs(rf["grad"]).max())
        out["loss_err"] = float(abs(g[-2] / n_glob - rf["loss"]))
        out["vgrad_err"] = float(np.abs(vbuf.numpy() - rvf["grad"]).max() / np.abs(rvf["grad"]).max())
        out["n_glob"] = n_glob
    dist.barrier()
    dist.destroy_process_group()


def test_dp_decomposition_world2_gloo():
    mgr = mp.get_context("spawn").Manager()  # fork() from a multi-threaded pytest process can deadlock
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["n_glob"] == 3000
    assert out["grad_err"] < 1e-5 and out["vgrad_err"] < 1e-5 and out["loss_err"] < 1e-6
    assert out["trpo_grad_err"] < 1e-5 and out["trpo_loss_err"] < 1e-6 and out["fvp_err"] < 1e-5 and out["kl_err"] < 1e-5
