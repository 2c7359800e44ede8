"""CPU: the C-ABI library loads without a GPU and exports every symbol include/b200rl.h declares; the ctypes
signature table covers them all; entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200rl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rl_[a-z0-9_]+)\s*\(", text)) - {"b200rl_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    from rl_replicas_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/b200rl.h but not exported by libb200rl.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)
    assert lib.b200rl_version() == 100


def test_struct_sizes_match_header_layout():
    from rl_replicas_b200 import _lib
    assert C.sizeof(_lib.MlpDesc) == 4 * 8
    assert C.sizeof(_lib.LossGradArgs) % 8 == 0
    assert C.sizeof(_lib.UpdateStats) == 5 * 8 + 4 * 4 + 4 * 8
    assert C.sizeof(_lib.PpoHparams) == 4 * 8 + 2 * 4 + 8 * 8 + 8
    assert C.sizeof(_lib.TrpoHparams) == 8 + 2 * 4 + 2 * 8
    assert C.sizeof(_lib.TrpoStats) == 5 * 8 + 4 * 4


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rl_replicas_b200 import _lib
    from rl_replicas_b200.engine import OnPolicyEngine
    with pytest.raises(_lib.B200RLError):
        OnPolicyEngine([17, 64, 64, 6], [17, 64, 64, 1], "gaussian", 1000, 10)
    lib = _lib.load()
    d = _lib.MlpDesc.make([17, 64, 64, 6])
    assert lib.b200rl_mlp_param_count(d) == 5702
    assert lib.b200rl_mlp_grid(d, 1000, 1) == -1  # no device: reported, not emulated


def test_host_logic_packing_and_descriptions():
    import numpy as np
    import torch
    from rl_replicas_b200 import synthetic
    from rl_replicas_b200.algorithms._onpolicy import describe_mlp, flat_params, write_flat
    from rl_replicas_b200.experience import Experience
    from rl_replicas_b200.networks import MLP
    from rl_replicas_b200.packing import pack_experience
    b = synthetic.ragged_batch(500, 4, 2, True, seed=0)
    exp = Experience(**synthetic.to_experience_lists(b, True))
    p = pack_experience(exp)
    for k in ("obs", "act", "rew", "last_obs", "ep_offsets"):
        np.testing.assert_array_equal(p[k], b[k])
    np.testing.assert_array_equal(p["ep_done"], b["ep_done"])
    net = MLP([4, 64, 32, 2], torch.nn.ReLU, torch.nn.Tanh)
    sizes, hid, out, lin = describe_mlp(net)
    assert sizes == [4, 64, 32, 2] and hid == "relu" and out == "tanh"
    f = flat_params(lin)
    np.testing.assert_array_equal(f, torch.nn.utils.parameters_to_vector(net.parameters()).detach().numpy())
    write_flat(lin, f * 2)
    np.testing.assert_allclose(flat_params(lin), f * 2)
    with pytest.raises(NotImplementedError):
        describe_mlp(torch.nn.Sequential(torch.nn.Conv1d(1, 1, 1)))
