"""The cfconv aggregate kernels alone, on the bench workload's graph (two stacked views of 256 x 30 atoms), with one filter
row per atom pair (shared) and per edge.  (Round 2 used this script to A/B a persistent-grid edition, see the negative
results noted in csrc/cfconv.cu; `persistent` is now always 0.)  Second part: the pair-centric kernel
(geossl_cfconv_pairs) over its tuning codes (warps per graph x filter rows in flight), checked against the row-gather result.
Back-to-back launches cycle over 4 filter tensors (4 x 112/225 MB > L2), CUDA events, mean of 40 launches.

    python profiles/bench_cfconv.py > profiles/rNN_vK_cfconv_ab.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geossl_b200 import ops  # noqa: E402
from geossl_b200.data import synthetic_batch  # noqa: E402

dev = "cuda:0"
b = synthetic_batch(256, 30, seed=0).to(dev)
pos = torch.cat([b.positions, b.positions + 0.3 * torch.randn_like(b.positions)])
g = ops.radius_csr(pos, torch.cat([b.batch, b.batch + 256]), 10.0, num_graphs=512)
g.ensure_pairs()
n, e, u = g.n_atoms, g.num_edges, int(g.n_pairs_dev.item())
print(f"atoms {n}, edges {e}, pairs {u}")
x = torch.randn(n, 128, device=dev)
filts = [torch.randn(g.capacity, 128, device=dev) for _ in range(4)]


def run(fn, reps=40):
    """Mean duration of ``fn`` over ``reps`` back-to-back launches replayed from a CUDA graph (the eager ctypes call path
    costs ~20 us of host time per launch, more than the faster kernels take)."""
    for i in range(4):
        fn(filts[i % 4])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(reps):
            fn(filts[i % 4])
    graph.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    graph.replay()
    t1.record()
    torch.cuda.synchronize()
    return 1e3 * t0.elapsed_time(t1) / reps


ref = {}
for persistent in (False,):
    for shared in (True, False):
        fr = g.pair_of_edge if shared else None
        rows = u if shared else e
        fwd = run(lambda f: ops._cfconv_fwd(x, f, g, fr))
        bwd = run(lambda f: ops._cfconv_bwd_x(f, x, g, fr))
        byt = 4 * 128 * rows + 2 * 4 * 128 * n + 4 * e * (2 if shared else 1) + 4 * (n + 1)
        out = (ops._cfconv_fwd(x, filts[0], g, fr), ops._cfconv_bwd_x(filts[0], x, g, fr))
        key = shared
        if key in ref:
            assert torch.equal(out[0], ref[key][0]) and torch.equal(out[1], ref[key][1]), "editions must agree bit for bit"
        ref[key] = out
        print(f"persistent={int(persistent)} shared={int(shared)}: fwd {fwd:6.1f} us ({byt / fwd / 1e3:6.0f} GB/s algorithmic), bwd_x {bwd:6.1f} us")

# ---- pair-centric kernel: every filter row crosses L2 -> SM once
g.max_graph_atoms = 30
fr = g.pair_of_edge
want = (ops._cfconv_fwd(x, filts[0], g, fr), ops._cfconv_bwd_x(filts[0], x, g, fr))
byt = 4 * 128 * u + 2 * 4 * 128 * n + 8 * u + 4 * (n + 1) + 4 * 513
for tuning in [int(t) for t in os.environ.get("GEOSSL_BENCH_CFCONV_TUNINGS", "116,208,216,308,316,408,416").split(",")]:
    ops.CFCONV_PAIRS_TUNING = tuning
    got = (ops._cfconv_pairs(x, filts[0], g, False), ops._cfconv_pairs(x, filts[0], g, True))
    err = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(got, want))
    again = ops._cfconv_pairs(x, filts[0], g, False)
    fwd = run(lambda f: ops._cfconv_pairs(x, f, g, False))
    bwd = run(lambda f: ops._cfconv_pairs(x, f, g, True))
    print(f"pairs tuning={tuning}: fwd {fwd:6.1f} us ({byt / fwd / 1e3:6.0f} GB/s algorithmic), bwd_x {bwd:6.1f} us, "
          f"max rel diff vs row-gather {err:.1e}, repeatable {torch.equal(again, got[0])}")

