"""The reference's public helper functions (ref: utils.py) with the same names, arguments and return types.

Inside ``train()`` their arithmetic is fused into the update engine (one segmented scan, statistics applied on
load ...); the functions here serve code that calls them directly -- custom algorithms, evaluation scripts.  They run
on the GPU through the C ABI (``b200rl_discounted_cumsum``, ``b200rl_gae_f64``, ``b200rl_mlp_loss_grad`` with the EVAL
loss, ``b200rl_normalize``, ``b200rl_polyak``); like the rest of the package there is no CPU fallback, a missing
library or device raises.  torch is used for device memory and the host<->device copies only."""
import ctypes as C
import random
from typing import Iterable, List

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib
from .policies import Policy


def _stream() -> int:
    from .engine import current_stream_handle
    return current_stream_handle()


def _p(t: Tensor):
    return C.c_void_p(t.data_ptr())


def discounted_cumulative_sums(vector: np.ndarray, discount: float) -> np.ndarray:
    """``y[t] = x[t] + discount * y[t+1]`` in float64 (ref: utils.py:14-28, scipy.signal.lfilter over the reversed
    vector).  Columns of a 2-D input are independent series (``axis=0``, as in the reference)."""
    lib = _lib.load()
    stream = _stream()
    x = np.asarray(vector, dtype=np.float64)
    if x.ndim == 0:
        raise ValueError("discounted_cumulative_sums needs at least a 1-D vector")
    cols = x.reshape(x.shape[0], -1).T  # [K, n]: one contiguous series per row
    out = np.empty_like(cols)
    for k in range(cols.shape[0]):
        d_x = torch.from_numpy(np.ascontiguousarray(cols[k])).cuda()
        d_y = torch.empty_like(d_x)
        _lib.check(lib.b200rl_discounted_cumsum(_p(d_x), d_x.numel(), float(discount), _p(d_y), stream),
                   "discounted_cumsum")
        out[k] = d_y.cpu().numpy()
    return out.T.reshape(x.shape)


def gae(rewards: np.ndarray, gamma: float, values: np.ndarray, gae_lambda: float) -> np.ndarray:
    """Generalised advantage estimates of ONE episode (ref: utils.py:31-44): ``rewards`` and ``values`` hold L+1
    entries (bootstrapped reward / V(last observation) at the end); returns L float64 advantages."""
    lib = _lib.load()
    r = np.ascontiguousarray(rewards, dtype=np.float64)
    v = np.asarray(values)
    f32 = v.dtype == np.float32  # what compute_values returns; numpy keeps gamma * values in float32 then
    v = np.ascontiguousarray(v, dtype=np.float32 if f32 else np.float64)
    if r.ndim != 1 or r.shape != v.shape or r.size < 1:
        raise ValueError("gae: rewards and values must be 1-D and of equal length (L + 1)")
    n = r.size - 1
    d_r, d_v = torch.from_numpy(r).cuda(), torch.from_numpy(v).cuda()
    d_out = torch.empty(max(n, 1), dtype=torch.float64, device="cuda")
    _lib.check(lib.b200rl_gae_f64(_p(d_r), _p(d_v), int(f32), n, float(gamma), float(gae_lambda), _p(d_out), _stream()),
               "gae_f64")
    return d_out[:n].cpu().numpy()


def polyak_average(params: Iterable[nn.Parameter], target_params: Iterable[nn.Parameter], rho: float) -> None:
    """``target <- rho * target + (1 - rho) * param`` for every pair, in place (ref: utils.py:47-57)."""
    lib = _lib.load()
    stream = _stream()
    with torch.no_grad():
        for param, target in zip(params, target_params):
            if param.shape != target.shape:
                raise ValueError("polyak_average: parameter shapes differ")
            d_t = target.data.detach().to(device="cuda", dtype=torch.float32).contiguous()
            d_p = param.data.detach().to(device="cuda", dtype=torch.float32).contiguous()
            d_work = d_t  # aliases target.data when that is a contiguous float32 CUDA tensor (updated in place)
            _lib.check(lib.b200rl_polyak(_p(d_work), _p(d_p), d_work.numel(), float(rho), stream), "polyak")
            if d_work.data_ptr() != target.data.data_ptr():
                target.data.copy_(d_work.to(target.device))


def compute_values(observations_with_last_observation: List[List[np.ndarray]], value_function) -> List[np.ndarray]:
    """V(obs) for every step of every episode plus its last observation (ref: utils.py:60-71): one float32 array of
    length L_e + 1 per episode.  All episodes go through ONE forward launch of the fused MLP kernel (EVAL loss)."""
    from .algorithms._onpolicy import describe_mlp, flat_params
    lib = _lib.load()
    if len(observations_with_last_observation) == 0:
        return []
    sizes, hidden, out_act, linears = describe_mlp(value_function.network)
    if out_act != "identity" or sizes[-1] != 1:
        raise NotImplementedError("compute_values: the value network must end in a single identity output")
    lengths = [len(ep) for ep in observations_with_last_observation]
    obs = np.concatenate([np.stack(ep).reshape(len(ep), -1) for ep in observations_with_last_observation])
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    if obs.shape[1] != sizes[0]:
        raise ValueError(f"compute_values: observations have {obs.shape[1]} features, the network expects {sizes[0]}")
    n = obs.shape[0]
    a = _lib.LossGradArgs()
    a.mlp = _lib.MlpDesc.make(sizes, hidden, "identity")
    a.loss, a.dist, a.n_rows, a.n_global = _lib.LOSS["eval"], _lib.DIST["none"], n, n
    d_obs = torch.from_numpy(obs).cuda()
    d_par = torch.from_numpy(flat_params(linears)).cuda()
    d_out = torch.empty(n, dtype=torch.float32, device="cuda")
    a.params, a.obs, a.row_out = d_par.data_ptr(), d_obs.data_ptr(), d_out.data_ptr()
    _lib.check(lib.b200rl_mlp_loss_grad(C.byref(a), _stream()), "mlp_loss_grad(eval)")
    flat = d_out.cpu().numpy()
    return [part.copy() for part in np.split(flat, np.cumsum(lengths)[:-1])]


def bootstrap_rewards_with_last_values(rewards: List[List[float]], episode_dones: List[bool],
                                       last_values: List[float]) -> List[np.ndarray]:
    """Per episode ``rewards + [0 if done else V(last observation)]`` (ref: utils.py:74-87).  Pure list
    bookkeeping (no arithmetic): inside train() it is folded into the scan's episode tail."""
    return [np.asarray(list(r) + [0 if done else last]) for r, done, last in zip(rewards, episode_dones, last_values)]


def normalize_tensor(vector: Tensor) -> Tensor:
    """``(v - mean(v)) / std(v)``, unbiased std, no epsilon (ref: utils.py:90-92)."""
    lib = _lib.load()
    flat = vector.detach().to(device="cuda", dtype=torch.float32).contiguous().reshape(-1)
    if flat.numel() == 0:
        raise ValueError("normalize_tensor: empty tensor")
    out = torch.empty_like(flat)
    _lib.check(lib.b200rl_normalize(_p(flat), flat.numel(), _p(out), _stream()), "normalize")
    return out.reshape(vector.shape).to(device=vector.device, dtype=vector.dtype)


class _NoisedPolicy(Policy):
    """Gaussian exploration noise around a deterministic policy, clipped to the action limit
    (ref: utils.py:101-124; numpy global RNG for get_action_numpy, torch RNG for get_action_tensor)."""

    def __init__(self, base_policy: Policy, action_space, action_noise_scale: float):
        super().__init__()
        self.base_policy = base_policy
        self.action_space = action_space
        self.action_noise_scale = action_noise_scale
        self.action_limit = action_space.high[0]
        self.action_size = action_space.shape[0]

    def get_action_tensor(self, observation: Tensor) -> Tensor:
        action = self.base_policy.get_action_tensor(observation)
        action += self.action_noise_scale * torch.randn(self.action_size)
        return torch.clip(action, -self.action_limit, self.action_limit)

    def get_action_numpy(self, observation: np.ndarray) -> np.ndarray:
        action = self.base_policy.get_action_numpy(observation)
        action += self.action_noise_scale * np.random.randn(self.action_size)
        return np.clip(action, -self.action_limit, self.action_limit)


def add_noise_to_get_action(policy: Policy, action_space, action_noise_scale: float) -> Policy:
    return _NoisedPolicy(policy, action_space, action_noise_scale)


def set_seed_for_libraries(seed: int) -> None:
    """Seed python / numpy / torch and make torch deterministic (ref: utils.py:127-136)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.use_deterministic_algorithms(True)
