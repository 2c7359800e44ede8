"""CPU oracle for the policy-gradient update path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
thing timed *as the CPU baseline*), never as a fallback for the CUDA path.

The oracle is a numpy restatement of the reference's algorithm
(rl_replicas 0.0.7, /root/reference @ 19d890dc).  Each function cites the
reference file:line it follows.  Third-party arithmetic on the path that is
NOT under /root/reference is restated from its published algorithm:

* torch 2.5.1 (uv.lock:764-765): nn.Linear / Tanh / ReLU, autograd,
  torch.optim.Adam (single-tensor CPU loop), torch.distributions
  {Categorical, Normal, Independent, kl_divergence}.
* scipy 1.14.1 (uv.lock:622-623): scipy.signal.lfilter (used as-is where
  the reference calls it; a plain loop restatement is pinned against it).

Pinning: the reference's own tests hold NO golden vector for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the
reference itself, run in the build container by ``tests/golden/make_golden.py``
(committed) and stored under ``tests/golden/*.npz``.
"""
