"""Python handle of the native on-policy update engine (C ABI: include/b200rl.h).

PyTorch is used for plumbing only: picking the CUDA device / current stream, exposing engine-owned device memory
as tensors (``view``) and torch.distributed collectives for data-parallel runs.  All arithmetic of the update path
runs in libb200rl.so's CUDA kernels; there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import DIST, MlpDesc, OnPolicyConfig, PpoHparams, TrpoHparams, TrpoStats, UpdateStats, check

POLICY, OLD_POLICY, VALUE = 0, 1, 2
_OPENED_IPC: Dict[bytes, int] = {}  # CUDA IPC handle -> device pointer of the mapping in this process


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap engine-owned device memory without copying."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


def current_stream_handle() -> int:
    import torch
    if not torch.cuda.is_available():
        raise _lib.B200RLError("no CUDA device: the B200 update engine has no CPU fallback")
    return int(torch.cuda.current_stream().cuda_stream)


class OnPolicyEngine:
    """Device-resident state of one PPO / VPG / TRPO learner."""

    def __init__(self, policy_sizes: Sequence[int], value_sizes: Sequence[int], dist: str, max_rows: int,
                 max_episodes: int, hidden_act: str = "tanh", rewards_f64: bool = True, train_log_std: bool = False):
        self.lib = _lib.load()
        current_stream_handle()  # fail early and loudly without a GPU
        cfg = OnPolicyConfig()
        cfg.policy = MlpDesc.make(policy_sizes, hidden_act, "identity")
        cfg.value = MlpDesc.make(value_sizes, hidden_act, "identity")
        cfg.dist = DIST[dist]
        cfg.rewards_f64 = int(rewards_f64)
        cfg.max_rows, cfg.max_episodes = int(max_rows), int(max_episodes)
        self.cfg = cfg
        self.dist = dist
        self.policy_sizes, self.value_sizes = list(policy_sizes), list(value_sizes)
        self.rewards_f64 = rewards_f64
        self.n_policy = int(self.lib.b200rl_mlp_param_count(cfg.policy))
        self.n_value = int(self.lib.b200rl_mlp_param_count(cfg.value))
        self.max_rows, self.max_episodes = int(max_rows), int(max_episodes)
        h = C.c_void_p()
        check(self.lib.b200rl_onpolicy_create(C.byref(cfg), C.byref(h)), "onpolicy_create")
        self.h = h
        self.train_log_std = bool(train_log_std)
        if self.train_log_std:  # the policy vector becomes [network parameters | log_std]
            check(self.lib.b200rl_onpolicy_set_train_log_std(h, 1), "set_train_log_std")
            self.n_policy += int(policy_sizes[-1])
        self.n_rows = 0
        self.n_episodes = 0
        self._allreduce_cb = None
        self._keep = []
        self.peer_exchange = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rl_onpolicy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters / optimiser state -------------------------------------------------------------------------
    def _n(self, which):
        return self.n_value if which == VALUE else self.n_policy

    def set_params(self, which: int, flat: np.ndarray):
        a = _c(flat, np.float32)
        self._keep.append(a)
        check(self.lib.b200rl_onpolicy_set_params(self.h, which, _ptr(a), a.size, current_stream_handle()), "set_params")

    def get_params(self, which: int) -> np.ndarray:
        out = np.empty(self._n(which), dtype=np.float32)
        check(self.lib.b200rl_onpolicy_get_params(self.h, which, _ptr(out), out.size, current_stream_handle()), "get_params")
        return out

    def set_adam(self, which: int, exp_avg: Optional[np.ndarray], exp_avg_sq: Optional[np.ndarray], step: int):
        m = None if exp_avg is None else _c(exp_avg, np.float32)
        v = None if exp_avg_sq is None else _c(exp_avg_sq, np.float32)
        self._keep += [m, v]
        check(self.lib.b200rl_onpolicy_set_adam(self.h, which, _ptr(m), _ptr(v), self._n(which), int(step),
                                                current_stream_handle()), "set_adam")

    def get_adam(self, which: int):
        m = np.empty(self._n(which), dtype=np.float32)
        v = np.empty(self._n(which), dtype=np.float32)
        step = C.c_int64()
        check(self.lib.b200rl_onpolicy_get_adam(self.h, which, _ptr(m), _ptr(v), m.size, C.byref(step),
                                                current_stream_handle()), "get_adam")
        return m, v, int(step.value)

    def set_log_std(self, log_std: np.ndarray):
        a = _c(log_std, np.float32)
        self._keep.append(a)
        check(self.lib.b200rl_onpolicy_set_log_std(self.h, _ptr(a), a.size, current_stream_handle()), "set_log_std")

    # ---- batch ----------------------------------------------------------------------------------------------
    def load_batch(self, batch: Dict[str, np.ndarray]):
        """Copy a packed HOST batch (see synthetic.py for the layout) to the device."""
        obs = _c(batch["obs"], np.float32)
        act = _c(batch["act"], np.float32)
        rew = _c(batch["rew"], np.float64 if self.rewards_f64 else np.float32)
        last = _c(batch["last_obs"], np.float32)
        off = _c(batch["ep_offsets"], np.int64)
        done = _c(batch["ep_done"], np.uint8)
        n, e = obs.shape[0], done.shape[0]
        if obs.ndim != 2 or obs.shape[1] != self.policy_sizes[0]:
            raise ValueError(f"observations must be [N,{self.policy_sizes[0]}], got {obs.shape}")
        want_act = (n, self.policy_sizes[-1]) if self.dist == "gaussian" else (n,)
        if act.shape != want_act:
            raise ValueError(f"actions must have shape {want_act}, got {act.shape}")
        if rew.shape != (n,) or last.shape != (e, obs.shape[1]) or off.shape != (e + 1,):
            raise ValueError("inconsistent packed batch shapes")
        self._keep += [obs, act, rew, last, off, done]  # keep host buffers alive until the stream has consumed them
        check(self.lib.b200rl_onpolicy_load_batch(self.h, _ptr(obs), _ptr(act), _ptr(rew), _ptr(last), _ptr(off),
                                                  _ptr(done), n, e, 0, current_stream_handle()), "load_batch")
        self.n_rows, self.n_episodes = n, e

    def load_batch_device(self, obs, act, rew, last_obs, ep_offsets, ep_done):
        """Same from torch CUDA tensors already resident in HBM (device-to-device copies)."""
        n, e = obs.shape[0], ep_done.shape[0]
        ts = [obs, act, rew, last_obs, ep_offsets, ep_done]
        assert all(t.is_cuda and t.is_contiguous() for t in ts)
        check(self.lib.b200rl_onpolicy_load_batch(self.h, *[C.c_void_p(t.data_ptr()) for t in ts], n, e, 1,
                                                  current_stream_handle()), "load_batch(device)")
        self.n_rows, self.n_episodes = n, e

    # ---- update ---------------------------------------------------------------------------------------------
    @staticmethod
    def hparams(gamma=0.99, gae_lambda=0.97, clip_range=0.2, max_kl_divergence=0.01, num_policy_gradients=80,
                num_value_gradients=80, policy_adam=(3e-4, 0.9, 0.999, 1e-8), value_adam=(1e-3, 0.9, 0.999, 1e-8),
                n_global_rows=0) -> PpoHparams:
        hp = PpoHparams()
        hp.gamma, hp.gae_lambda, hp.clip_range, hp.max_kl_divergence = gamma, gae_lambda, clip_range, max_kl_divergence
        hp.num_policy_gradients, hp.num_value_gradients = int(num_policy_gradients), int(num_value_gradients)
        hp.policy_lr, hp.policy_beta1, hp.policy_beta2, hp.policy_eps = policy_adam
        hp.value_lr, hp.value_beta1, hp.value_beta2, hp.value_eps = value_adam
        hp.n_global_rows = int(n_global_rows)
        return hp

    def _make_allreduce(self, process_group):
        import torch
        import torch.distributed as dist

        def cb(user, buf, count, dtype, stream):
            try:
                t = torch.as_tensor(_CudaArray(buf, count, "<f8" if dtype == 1 else "<f4"), device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
                return 0
            except Exception as exc:  # surfaced by the engine as a failed update
                import traceback
                traceback.print_exc()
                return 1

        return _lib.ALLREDUCE_FN(cb)

    def _make_allreduce_from(self, fn):
        import torch

        def cb(user, buf, count, dtype, stream):
            try:
                fn(torch.as_tensor(_CudaArray(buf, count, "<f8" if dtype == 1 else "<f4"), device="cuda"))
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return 1

        return _lib.ALLREDUCE_FN(cb)

    # ---- one-shot gradient exchange over peer-mapped memory (one node, NVLink) ---------------------------------
    def comm_export(self):
        """(64-byte CUDA IPC handle, local device pointer) of this engine's exchange buffer."""
        handle = (C.c_uint8 * 64)()
        ptr = C.c_void_p()
        check(self.lib.b200rl_onpolicy_comm_export(self.h, handle, C.byref(ptr)), "comm_export")
        return bytes(handle), int(ptr.value)

    def comm_attach(self, rank: int, peer_ptrs: Sequence[int]):
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        check(self.lib.b200rl_onpolicy_comm_attach(self.h, int(rank), len(peer_ptrs), arr), "comm_attach")
        self.peer_exchange = True

    def enable_peer_exchange(self, process_group=None) -> bool:
        """Exchange IPC handles over torch.distributed and map every rank's buffer (all ranks on ONE node).  Returns
        False (and leaves the NCCL all-reduce in place) when the ranks cannot map each other's memory."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
        if world < 2 or world > 16:
            return False
        ok = 1
        try:
            handle, mine = self.comm_export()
        except _lib.B200RLError:
            handle, mine, ok = b"", 0, 0
        gathered = [None] * world
        dist.all_gather_object(gathered, (handle, ok), group=process_group)
        ptrs = []
        if all(g[1] for g in gathered):
            # mappings of buffers that no longer exist on their rank (its engine was replaced) go first: a new
            # allocation may land where the old one was, and that cannot be mapped twice
            live = {hd for r, (hd, _) in enumerate(gathered) if r != rank}
            for hd in [h_ for h_ in _OPENED_IPC if h_ not in live]:
                self.lib.b200rl_ipc_close(C.c_void_p(_OPENED_IPC.pop(hd)))
            for r, (hd, _) in enumerate(gathered):
                if r == rank:
                    ptrs.append(mine)
                    continue
                if hd not in _OPENED_IPC:  # a handle can be mapped once per process: keep the mapping
                    p = C.c_void_p()
                    buf = (C.c_uint8 * 64).from_buffer_copy(hd)
                    if self.lib.b200rl_ipc_open(buf, C.byref(p)) != 0:
                        ok = 0
                        break
                    _OPENED_IPC[hd] = int(p.value)
                ptrs.append(_OPENED_IPC[hd])
        else:
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=process_group)  # all or nobody
        if int(t.item()) != 1:
            return False
        self.comm_attach(rank, ptrs)
        return True

    def update(self, hp: PpoHparams, algo: str = "ppo", process_group=None, distributed: bool = False,
               allreduce=None) -> UpdateStats:
        """``distributed``: all-reduce through torch.distributed on ``process_group`` (NCCL).  ``allreduce``: a callable
        ``fn(tensor)`` that sums the given CUDA tensor over the data-parallel group in place, for transports other than
        torch.distributed (and for tests)."""
        stats = UpdateStats()
        cb = None
        if allreduce is not None:
            self._allreduce_cb = self._make_allreduce_from(allreduce)
            cb = C.cast(self._allreduce_cb, C.c_void_p)
        elif distributed:
            self._allreduce_cb = self._make_allreduce(process_group)
            cb = C.cast(self._allreduce_cb, C.c_void_p)
        fn = {"ppo": self.lib.b200rl_ppo_update, "vpg": self.lib.b200rl_vpg_update}[algo]
        check(fn(self.h, C.byref(hp), cb, None, C.byref(stats), current_stream_handle()), f"{algo}_update")
        self._keep.clear()  # the update synchronised the stream: staged host buffers are no longer in flight
        return stats

    def trpo_update(self, hp: PpoHparams, max_constraint=0.01, n_conjugate_gradients=10, max_backtracks=15,
                    backtrack_ratio=0.8, hvp_damping_coefficient=1e-5, process_group=None, distributed: bool = False,
                    allreduce=None):
        """The reference's TRPO.train on the loaded batch; returns (UpdateStats, TrpoStats).  ``distributed`` /
        ``allreduce`` as in ``update``: every rank holds a block of episodes, ``hp.n_global_rows`` the global row count,
        and every batch-derived sum of the step is all-reduced (b200rl_trpo_update_dp)."""
        cg = TrpoHparams()
        cg.max_constraint, cg.n_conjugate_gradients = float(max_constraint), int(n_conjugate_gradients)
        cg.max_backtracks, cg.backtrack_ratio = int(max_backtracks), float(backtrack_ratio)
        cg.hvp_damping_coefficient = float(hvp_damping_coefficient)
        stats, ts = UpdateStats(), TrpoStats()
        cb = None
        if allreduce is not None:
            self._allreduce_cb = self._make_allreduce_from(allreduce)
            cb = C.cast(self._allreduce_cb, C.c_void_p)
        elif distributed:
            self._allreduce_cb = self._make_allreduce(process_group)
            cb = C.cast(self._allreduce_cb, C.c_void_p)
        check(self.lib.b200rl_trpo_update_dp(self.h, C.byref(hp), C.byref(cg), cb, None, C.byref(stats), C.byref(ts),
                                             current_stream_handle()), "trpo_update")
        self._keep.clear()
        return stats, ts

    def fvp(self, v: np.ndarray, damping: float = 1e-5) -> np.ndarray:
        """(F + damping I) v at the current policy parameters on the loaded batch (one fused FVP launch)."""
        v = _c(v, np.float32)
        out = np.empty_like(v)
        check(self.lib.b200rl_onpolicy_fvp(self.h, _ptr(v), _ptr(out), v.size, float(damping), current_stream_handle()),
              "fvp")
        return out

    def scalar_history(self) -> np.ndarray:
        """[slots, 8] float64 scalar sums of the last update (see b200rl_onpolicy_scalar_history)."""
        n = C.c_int32()
        probe = np.zeros((1, _lib.N_SCALARS))
        check(self.lib.b200rl_onpolicy_scalar_history(self.h, _ptr(probe), 0, C.byref(n)), "scalar_history")
        out = np.zeros((max(int(n.value), 1), _lib.N_SCALARS))
        check(self.lib.b200rl_onpolicy_scalar_history(self.h, _ptr(out), int(n.value), C.byref(n)), "scalar_history")
        return out[:int(n.value)]

    def run_stage(self, stage: str, hp: PpoHparams):
        check(self.lib.b200rl_onpolicy_run_stage(self.h, stage.encode(), C.byref(hp), current_stream_handle()),
              f"run_stage({stage})")

    def view(self, name: str):
        """Engine-owned device buffer as a torch tensor (zero copy)."""
        import torch
        p, n, d = C.c_void_p(), C.c_int64(), C.c_int32()
        check(self.lib.b200rl_onpolicy_device_view(self.h, name.encode(), C.byref(p), C.byref(n), C.byref(d)), "device_view")
        return torch.as_tensor(_CudaArray(p.value, int(n.value), "<f8" if d.value == 1 else "<f4"), device="cuda")


class OffPolicyEngine:
    """Device-resident state of one DDPG / TD3 learner (C ABI: b200rl_offpolicy_*)."""

    NETS = {"policy": 0, "q1": 1, "q2": 2, "target_policy": 3, "target_q1": 4, "target_q2": 5}

    def __init__(self, policy_sizes, q_sizes, n_q: int, max_minibatch: int, max_steps: int, policy_acts=("relu", "tanh"),
                 q_acts=("relu", "identity")):
        from ._lib import OffPolicyConfig
        self.lib = _lib.load()
        current_stream_handle()
        cfg = OffPolicyConfig()
        cfg.policy = MlpDesc.make(policy_sizes, *policy_acts)
        cfg.q = MlpDesc.make(q_sizes, *q_acts)
        cfg.n_q, cfg.max_minibatch, cfg.max_steps = int(n_q), int(max_minibatch), int(max_steps)
        self.n_q, self.max_minibatch, self.max_steps = int(n_q), int(max_minibatch), int(max_steps)
        self.policy_sizes, self.q_sizes = list(policy_sizes), list(q_sizes)
        self.policy_acts, self.q_acts = tuple(policy_acts), tuple(q_acts)
        self.n_policy = int(self.lib.b200rl_mlp_param_count(cfg.policy))
        self.n_qp = int(self.lib.b200rl_mlp_param_count(cfg.q))
        h = C.c_void_p()
        check(self.lib.b200rl_offpolicy_create(C.byref(cfg), C.byref(h)), "offpolicy_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rl_offpolicy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _n(self, which):
        return self.n_policy if which in (0, 3) else self.n_qp

    def set_params(self, which: int, flat: np.ndarray):
        a = _c(flat, np.float32)
        check(self.lib.b200rl_offpolicy_set_params(self.h, which, _ptr(a), a.size, current_stream_handle()), "set_params")
        import torch
        torch.cuda.current_stream().synchronize()  # `a` may be a temporary

    def get_params(self, which: int) -> np.ndarray:
        out = np.empty(self._n(which), dtype=np.float32)
        check(self.lib.b200rl_offpolicy_get_params(self.h, which, _ptr(out), out.size, current_stream_handle()), "get_params")
        return out

    def set_adam(self, which: int, m, v, step: int):
        m = None if m is None else _c(m, np.float32)
        v = None if v is None else _c(v, np.float32)
        check(self.lib.b200rl_offpolicy_set_adam(self.h, which, _ptr(m), _ptr(v), self._n(which), int(step),
                                                 current_stream_handle()), "set_adam")
        import torch
        torch.cuda.current_stream().synchronize()

    def get_adam(self, which: int):
        m, v = np.empty(self._n(which), np.float32), np.empty(self._n(which), np.float32)
        step = C.c_int64()
        check(self.lib.b200rl_offpolicy_get_adam(self.h, which, _ptr(m), _ptr(v), m.size, C.byref(step),
                                                 current_stream_handle()), "get_adam")
        return m, v, int(step.value)

    # ---- whole state in one transfer ----
    def state_layout(self):
        """[(kind, net index, offset, count)] of the state blob: ("params", 0..5) then ("m" / "v", 0..2); every
        segment starts on a multiple of 64 floats (b200rl.h)."""
        pad = lambda n: (n + 63) & ~63
        present = [0, 1] + ([2] if self.n_q == 2 else []) + [3, 4] + ([5] if self.n_q == 2 else [])
        out, off = [], 0
        for i in present:
            out.append(("params", i, off, self._n(i)))
            off += pad(self._n(i))
        for i in [0, 1] + ([2] if self.n_q == 2 else []):
            for kind in ("m", "v"):
                out.append((kind, i, off, self._n(i)))
                off += pad(self._n(i))
        assert off == int(self.lib.b200rl_offpolicy_state_floats(self.h))
        return out, off

    def state_buffer(self):
        """The engine's persistent host-side state blob (a float32 torch tensor, page-locked when CUDA allows it): filled
        by ``get_state()``, sent by ``set_state()``.  Views into it stay valid for the engine's lifetime."""
        import torch
        buf = getattr(self, "_state_buf", None)
        if buf is None:
            n = int(self.lib.b200rl_offpolicy_state_floats(self.h))
            buf = torch.zeros(n, dtype=torch.float32)
            try:
                buf = buf.pin_memory()
            except RuntimeError:  # pragma: no cover  (no page-locked memory available)
                pass
            self._state_buf = buf
        return buf

    def get_state(self):
        """Device -> the persistent blob; returns (blob as numpy view, [3 Adam step counts])."""
        buf = self.state_buffer()
        steps = (C.c_int64 * 3)()
        check(self.lib.b200rl_offpolicy_get_state(self.h, C.c_void_p(buf.data_ptr()), buf.numel(), steps,
                                                  current_stream_handle()), "get_state")
        return buf.numpy(), [int(x) for x in steps]

    def set_state(self, blob, steps):
        """``blob`` = None sends the persistent blob (fill ``state_buffer()`` first), else any float32 array of that size."""
        buf = self.state_buffer()
        if blob is not None and not (isinstance(blob, np.ndarray) and blob.ctypes.data == buf.data_ptr()):
            buf.numpy()[:] = np.asarray(blob, dtype=np.float32).reshape(-1)
        st = (C.c_int64 * 3)(*[int(x) for x in steps])
        check(self.lib.b200rl_offpolicy_set_state(self.h, C.c_void_p(buf.data_ptr()), buf.numel(), st,
                                                  current_stream_handle()), "set_state")

    def train(self, hp, obs, act, rew, next_obs, done, noise=None):
        """obs/next_obs [S,B,O], act [S,B,A], rew/done [S,B], noise [S,B,A] or None -> dict of logged quantities."""
        obs, act, next_obs = _c(obs, np.float32), _c(act, np.float32), _c(next_obs, np.float32)
        rew, done = _c(rew, np.float32), _c(done, np.float32)
        noise = None if noise is None else _c(noise, np.float32)
        S, B = obs.shape[0], obs.shape[1]
        q1v, q2v = np.zeros((S, B), np.float32), np.zeros((S, B), np.float32)
        l1, l2, lp = np.zeros(S, np.float32), np.zeros(S, np.float32), np.zeros(max(S, 1), np.float32)
        npol = C.c_int32()
        check(self.lib.b200rl_offpolicy_train(self.h, C.byref(hp), S, B, _ptr(obs), _ptr(act), _ptr(rew), _ptr(next_obs),
                                              _ptr(done), _ptr(noise), _ptr(q1v), _ptr(q2v), _ptr(l1), _ptr(l2), _ptr(lp),
                                              C.byref(npol), current_stream_handle()), "offpolicy_train")
        return dict(q1_values=q1v, q2_values=q2v, q1_losses=l1, q2_losses=l2, policy_losses=lp[:npol.value])

    def train_gather_rng(self, hp, columns, rows: int, ring_start: int, ring_size: int, S: int, B: int, seed: int, call: int):
        """``train_gather`` with the indices and the smoothing noise drawn on the device (Philox keyed by ``seed``, block
        ``call``); ``ring_size`` live rows, logical row u at physical ``(ring_start + u) % rows``.  Opt-in: the streams
        are not the reference's numpy / torch ones."""
        q1v, q2v = np.zeros((S, B), np.float32), np.zeros((S, B), np.float32)
        l1, l2, lp = np.zeros(S, np.float32), np.zeros(S, np.float32), np.zeros(max(S, 1), np.float32)
        npol = C.c_int32()
        ptrs = [C.c_void_p(t.data_ptr()) for t in columns]
        check(self.lib.b200rl_offpolicy_train_gather_rng(self.h, C.byref(hp), S, B, *ptrs, int(rows), int(ring_start),
                                                         int(ring_size), int(seed) & (2 ** 64 - 1), int(call), _ptr(q1v),
                                                         _ptr(q2v), _ptr(l1), _ptr(l2), _ptr(lp), C.byref(npol),
                                                         current_stream_handle()), "offpolicy_train_gather_rng")
        return dict(q1_values=q1v, q2_values=q2v, q1_losses=l1, q2_losses=l2, policy_losses=lp[:npol.value])

    def get_draws(self, S: int, B: int, with_noise: bool = True):
        """(physical rows [S,B] int64, noise [S,B,A] float32 or None) of the last train_gather / train_gather_rng call."""
        idx = np.empty((S, B), np.int64)
        A = self.policy_sizes[-1]
        noise = np.empty((S, B, A), np.float32) if with_noise else None
        check(self.lib.b200rl_offpolicy_get_draws(self.h, S, B, _ptr(idx), _ptr(noise), current_stream_handle()), "get_draws")
        return idx, noise

    def train_gather(self, hp, columns, rows: int, idx, noise=None):
        """Minibatches gathered on the device: ``columns`` = CUDA float32 tensors (obs [rows,O], act [rows,A], rew [rows],
        next_obs [rows,O], done [rows]) of a device-resident replay buffer, ``idx`` [S,B] int64 physical rows (host)."""
        idx = _c(idx, np.int64)
        noise = None if noise is None else _c(noise, np.float32)
        S, B = idx.shape
        q1v, q2v = np.zeros((S, B), np.float32), np.zeros((S, B), np.float32)
        l1, l2, lp = np.zeros(S, np.float32), np.zeros(S, np.float32), np.zeros(max(S, 1), np.float32)
        npol = C.c_int32()
        ptrs = [C.c_void_p(t.data_ptr()) for t in columns]
        check(self.lib.b200rl_offpolicy_train_gather(self.h, C.byref(hp), S, B, *ptrs, int(rows), _ptr(idx), _ptr(noise),
                                                     _ptr(q1v), _ptr(q2v), _ptr(l1), _ptr(l2), _ptr(lp), C.byref(npol),
                                                     current_stream_handle()), "offpolicy_train_gather")
        return dict(q1_values=q1v, q2_values=q2v, q1_losses=l1, q2_losses=l2, policy_losses=lp[:npol.value])
