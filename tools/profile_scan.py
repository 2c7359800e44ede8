"""Run the GAE scan on the 65536 x 1000 shape a few times (for ncu captures)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rl_replicas_b200 import _lib  # noqa: E402

lib = _lib.load()
E2 = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T2 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n2 = E2 * T2
rew = torch.randn(n2, dtype=torch.float64, device="cuda")
val = torch.randn(n2, dtype=torch.float32, device="cuda")
lv = torch.randn(E2, dtype=torch.float32, device="cuda")
off = torch.arange(E2 + 1, dtype=torch.int64, device="cuda") * T2
done = (torch.rand(E2, device="cuda") < 0.9).to(torch.uint8)
adv, ret = torch.empty(n2, dtype=torch.float32, device="cuda"), torch.empty(n2, dtype=torch.float32, device="cuda")
st = torch.zeros(3, dtype=torch.float64, device="cuda")
wsb = lib.b200rl_gae_scan_workspace_bytes(n2)
ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3):
    _lib.check(lib.b200rl_gae_scan(p(rew), 1, p(val), p(lv), p(off), p(done), n2, E2, 0.99, 0.97, p(adv), p(ret), p(st),
                                   p(ws), wsb, int(torch.cuda.current_stream().cuda_stream)), "gae_scan")
torch.cuda.synchronize()
print("stats", st.cpu().numpy())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = int(torch.cuda.current_stream().cuda_stream)
a.record()
for _ in range(20):
    lib.b200rl_gae_scan(p(rew), 1, p(val), p(lv), p(off), p(done), n2, E2, 0.99, 0.97, p(adv), p(ret), p(st), p(ws), wsb, stream)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"E={E2} T={T2}: {ms:.4f} ms/launch, {20.0 * n2 / ms / 1e6:.1f} GB/s algorithmic")
