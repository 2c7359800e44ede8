"""Data-parallel parity check, run under torchrun on N GPUs (NCCL):
   torchrun --nproc-per-node N tests/dist_check.py
Every rank builds the same global batch, trains on its contiguous block of episodes with distributed=True, and the
result must match the numpy ORACLE on the whole batch at 1e-5 of max|ref| (north_star; see the note on PPO's clip
mask below), and a single-GPU run of the same engine at 1e-5: the sharded engine computes the same global
mean / std / KL and applies the same reduced gradient on every rank.  Both gradient-exchange
paths are exercised: the one-shot exchange over peer-mapped memory (default) and the NCCL all-reduce per iteration
(B200RL_PEER_EXCHANGE=0).  A data-parallel TRPO step is checked against the single-GPU one at the end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import onpolicy as O  # noqa: E402
from rl_replicas_b200 import synthetic  # noqa: E402
from test_gpu_ppo import build, flat  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(0)
    ps, vs = [17, 64, 64, 6], [17, 64, 64, 1]
    mk = lambda sz: [(rng.standard_normal((o, i)).astype(np.float32) / np.sqrt(i), np.zeros(o, np.float32))
                     for i, o in zip(sz[:-1], sz[1:])]
    pl, vl = mk(ps), mk(vs)
    log_std = np.full(6, -0.5, np.float32)
    full = synthetic.ragged_batch(40000, 17, 6, False, seed=3, min_len=50, max_len=400,
                                  mean_fn=lambda o: O.mlp_forward(pl, o)[0])
    hp = dict(num_policy_gradients=6, num_value_gradients=6)
    ok = True
    oracle = {}
    # (max_kl, clip range, peer exchange).  max_kl = 0.002 triggers the device-side early stop on every rank.
    # clip = 1e9 never clips: the surrogate is smooth in the parameters and six steps must agree with the oracle at 1e-5.
    # With the default clip = 0.2 a sample whose ratio sits at 1 +- 0.2 flips its mask on a one-ulp difference in the
    # parameters (it does so between any two float32 implementations, the reference's CPU arithmetic included), so
    # that configuration is held to 1e-5 against the SAME engine on one GPU (identical arithmetic, different
    # reduction tree) and to 1e-4 against the oracle.
    # peer: "1" publish + wait + gather + Adam in one launch per iteration, "3" the same as three launches, "0" NCCL
    cases = [(float("inf"), 1e9, "1"), (float("inf"), 0.2, "1"), (0.002, 0.2, "1"),
             (float("inf"), 1e9, "3"), (0.002, 0.2, "3"),
             (float("inf"), 1e9, "0"), (float("inf"), 0.2, "0"), (0.002, 0.2, "0")]
    for max_kl, clip, peer in cases:
        os.environ["B200RL_PEER_EXCHANGE"] = "0" if peer == "0" else "1"
        os.environ["B200RL_PEER_ONE_LAUNCH"] = "0" if peer == "3" else "1"
        dp = build(ps, vs, "gaussian", O.flatten_layers(pl), O.flatten_layers(vl), log_std, distributed=True,
                   max_kl_divergence=max_kl, clip_range=clip, **hp)
        dp.train_packed(synthetic.shard_batch(full, rank, world))
        path = ("peer-memory exchange, " + ("one launch" if peer == "1" else "three launches")) if dp._engine.peer_exchange \
            else "NCCL all-reduce"
        p_dp, v_dp = flat(dp.policy.network).copy(), flat(dp.value_function.network).copy()
        st = dp.last_update_stats
        # all ranks must hold bit-identical parameters
        t = torch.from_numpy(np.concatenate([p_dp, v_dp])).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        if rank == 0:
            ref = build(ps, vs, "gaussian", O.flatten_layers(pl), O.flatten_layers(vl), log_std,
                        max_kl_divergence=max_kl, clip_range=clip, **hp)
            ref.train_packed(full)
            rs = ref.last_update_stats
            ep = np.abs(p_dp - flat(ref.policy.network)).max() / np.abs(flat(ref.policy.network)).max()
            ev = np.abs(v_dp - flat(ref.value_function.network)).max() / np.abs(flat(ref.value_function.network)).max()
            key = (max_kl, clip)
            if key not in oracle:
                # gradient sums in float64: a float32 BLAS sum has noise of its own that Adam's g / (|g| + eps)
                # amplifies for near-zero entries (it depends on the host's BLAS kernels, not on the GPU result)
                oracle[key] = O.ppo_train(full, pl, vl, "gaussian", log_std, O.AdamState(5702, 3e-4),
                                          O.AdamState(5377, 1e-3), max_kl=max_kl, clip=clip, n_policy=6, n_value=6,
                                          acc=np.float64)
            o = oracle[key]
            if key + (32,) not in oracle:  # for the record: the same oracle with float32 gradient sums
                oracle[key + (32,)] = O.ppo_train(full, pl, vl, "gaussian", log_std, O.AdamState(5702, 3e-4),
                                                  O.AdamState(5377, 1e-3), max_kl=max_kl, clip=clip, n_policy=6, n_value=6)
            o32 = oracle[key + (32,)]
            e32 = np.abs(p_dp - o32["policy_flat"]).max() / np.abs(o32["policy_flat"]).max()
            eop = np.abs(p_dp - o["policy_flat"]).max() / np.abs(o["policy_flat"]).max()
            eov = np.abs(v_dp - o["value_flat"]).max() / np.abs(o["value_flat"]).max()
            bar_policy = 1e-5 if (clip > 1e6 or o["policy_steps"] <= 3) else 1e-4
            good = (same and ep < 1e-5 and ev < 1e-5 and eop < bar_policy and eov < 1e-5
                    and st.policy_steps_applied == rs.policy_steps_applied == o["policy_steps"]
                    and dp._engine.peer_exchange == (peer != "0")
                    and abs(st.kl_divergence - rs.kl_divergence) < 1e-4 * abs(rs.kl_divergence) + 1e-8
                    and abs(st.value_loss_mean - rs.value_loss_mean) < 1e-5 * rs.value_loss_mean
                    and abs(st.adv_std - rs.adv_std) < 1e-9 * rs.adv_std)
            print(f"max_kl={max_kl} clip={clip:g} [{path}]: world={world} ranks_identical={same} vs oracle: policy {eop:.2e} (float32-sum oracle {e32:.2e}) value {eov:.2e}; "
                  f"vs 1-GPU run: policy_err={ep:.2e} value_err={ev:.2e} "
                  f"steps {st.policy_steps_applied}/{rs.policy_steps_applied} kl {st.kl_divergence:.6g}/{rs.kl_divergence:.6g} "
                  f"vloss {st.value_loss_mean:.6g}/{rs.value_loss_mean:.6g} -> {'OK' if good else 'FAIL'}")
            ok = ok and good
    # ---- TRPO: the constrained step with every batch-derived sum all-reduced (b200rl_trpo_update_dp) ----
    from test_gpu_trpo import build_trpo
    g = dict(policy_sizes=np.asarray(ps), value_sizes=np.asarray(vs), policy_flat0=O.flatten_layers(pl),
             value_flat0=O.flatten_layers(vl), log_std=log_std)
    dp = build_trpo(g, num_value_gradients=4, distributed=True)
    dp.train_packed(synthetic.shard_batch(full, rank, world))
    p_dp, v_dp = flat(dp.policy.network).copy(), flat(dp.value_function.network).copy()
    t = torch.from_numpy(np.concatenate([p_dp, v_dp])).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    if rank == 0:
        ref = build_trpo(g, num_value_gradients=4)
        ref.train_packed(full)
        ts_d, ts_r = dp.last_trpo_stats, ref.last_trpo_stats
        ep = np.abs(p_dp - flat(ref.policy.network)).max() / np.abs(flat(ref.policy.network)).max()
        ev = np.abs(v_dp - flat(ref.value_function.network)).max() / np.abs(flat(ref.value_function.network)).max()
        # ten CG iterations amplify the different summation order of the reduced vectors: 2e-3 on the step (the bar of
        # tests/test_gpu_trpo.py), decisions exact
        good = (same and ep < 2e-3 and ev < 2e-5 and ts_d.accepted_index == ts_r.accepted_index
                and ts_d.rejected == ts_r.rejected == 0 and ts_d.fvp_launches == ts_r.fvp_launches)
        print(f"TRPO [NCCL all-reduce per reduced quantity]: world={world} ranks_identical={same} vs 1-GPU run: "
              f"policy_err={ep:.2e} value_err={ev:.2e} accepted {ts_d.accepted_index}/{ts_r.accepted_index} "
              f"kl {ts_d.kl:.6g}/{ts_r.kl:.6g} -> {'OK' if good else 'FAIL'}")
        ok = ok and good
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL")
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
