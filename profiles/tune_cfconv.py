"""A/B timing of the cfconv gather kernel variants on the bench workload (stacked 2 x 256 molecules x 30 atoms).
CUDA-event timing, L2 flushed between launches.  Not a bench number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geossl_b200 import _lib, ops
from geossl_b200.data import synthetic_batch

dev = "cuda:0"
lib = _lib.load()
b = synthetic_batch(512, 30, seed=0, with_pairs=False).to(dev)
g = ops.radius_csr(b.positions, b.batch, 10.0, num_graphs=512).ensure_transpose().ensure_pairs()
n = b.positions.shape[0]
xs = [torch.randn(n, 128, device=dev) for _ in range(4)]
filts = [torch.randn(g.capacity, 128, device=dev) for _ in range(4)]       # 4 x 225 MB: every launch misses L2


def timeit(fn, reps=40):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    c.record()
    torch.cuda.synchronize()
    return a.elapsed_time(c) * 1e3 / reps


print(f"atoms {n}, edges {g.num_edges}, pairs {int(g.n_pairs_dev.item())}")
ref = {}
for variant in (7, 3, 0):                                   # every variant must give bit-identical results (same edge order)
    lib.geossl_debug_set_cfconv_variant(variant | (variant << 3))
    o1, o2 = ops._cfconv_fwd(xs[0], filts[0], g, g.pair_of_edge), ops._cfconv_bwd_x(filts[0], xs[0], g, g.pair_of_edge)
    torch.cuda.synchronize()
    if not ref:
        ref = {"f": o1, "b": o2}
    else:
        assert torch.equal(o1, ref["f"]) and torch.equal(o2, ref["b"]), variant
print("variants 7 / 3 / 0 agree bit for bit")
for variant in (7, 3):
    lib.geossl_debug_set_cfconv_variant(variant | (variant << 3))
    for shared in (True, False):
        row = g.pair_of_edge if shared else None
        f = timeit(lambda i: ops._cfconv_fwd(xs[i % 4], filts[i % 4], g, row))
        bx = timeit(lambda i: ops._cfconv_bwd_x(filts[i % 4], xs[i % 4], g, row))
        print(f"variant {variant} shared={shared}: fwd {f:.1f} us, bwd_x {bx:.1f} us")
lib.geossl_debug_set_cfconv_variant(3 | (3 << 3))
