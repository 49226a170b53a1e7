"""Per-phase clock trace of CTA 0 of the tensor-core filter forward kernel (debug aid, not a bench)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geossl_b200 import _lib, ops  # noqa: E402
from geossl_b200.data import synthetic_batch  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
b = synthetic_batch(256, 30, seed=0, with_pairs=False).to(dev)
graph = ops.radius_csr(b.positions, b.batch, 10.0, num_graphs=256)
G = 50
w1, b1 = torch.randn(128, G, device=dev) * 0.3, torch.randn(128, device=dev) * 0.1
w2, b2 = torch.randn(128, 128, device=dev) * 0.15, torch.randn(128, device=dev) * 0.1
offset = torch.linspace(0, 10, G, device=dev)
for _ in range(3):
    ops.filter_forward(graph, offset, -12.0, 10.0, w1, b1, w2, b2, mode="tc_fp16")
buf = torch.zeros(512, dtype=torch.int64, device=dev)
_lib.check(lib.geossl_debug_set_trace(ctypes.c_void_p(buf.data_ptr())))
ops.filter_forward(graph, offset, -12.0, 10.0, w1, b1, w2, b2, mode="tc_fp16")
torch.cuda.synchronize()
_lib.check(lib.geossl_debug_set_trace(None))
t = buf.cpu().view(32, 16)
t0 = int(t[0][t[0] > 0].min())
names = ["P:phi_empty", "P:phi_done", "M:mma1_go", "M:mma1_issued", "M:mma2_go", "M:mma2_issued", "E:d1_full", "E:d1_read",
         "E:ssp_done", "E:s_empty", "E:s_stored", "E:d2_full", "E:e2_done"]
print("tile " + " ".join(f"{n:>13s}" for n in names))
for i in range(14):
    print(f"{i:4d} " + " ".join(f"{(int(t[i][k]) - t0) if t[i][k] > 0 else -1:13d}" for k in range(13)))

# ---- backward kernel
x = torch.randn(b.positions.shape[0], 128, device=dev, requires_grad=True)
leaves = [t.clone().requires_grad_() for t in (w1, b1, w2, b2)]
gout = torch.randn(b.positions.shape[0], 128, device=dev)
for it in range(4):
    if it == 3:
        buf.zero_()
        _lib.check(lib.geossl_debug_set_trace_bwd(ctypes.c_void_p(buf.data_ptr())))
    out = ops.CFConvLayer.apply(x, *leaves, offset, graph, -12.0, 10.0)
    out.backward(gout)
torch.cuda.synchronize()
_lib.check(lib.geossl_debug_set_trace_bwd(None))
t = buf.cpu().view(32, 16)
t0 = int(t[0][t[0] > 0].min())
names = ["P:phi_free", "P:phi_done", "P:du_free", "P:du_done", "M:mma3_go", "M:wg2_go", "M:wg2_issd", "M:wg1_go", "M:wg1_issd",
         "E:d1_full", "E:s_stored", "E:d3_full", "E:s_free", "E:da_stored"]
print("bwd  " + " ".join(f"{n:>11s}" for n in names))
for i in range(24):
    print(f"{i:4d} " + " ".join(f"{(int(t[i][k]) - t0) if t[i][k] > 0 else -1:11d}" for k in range(14)))
