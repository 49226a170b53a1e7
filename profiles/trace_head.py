"""Per-phase clock trace of CTA 0 of the tensor-core DDM head kernels (debug aid, not a bench)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geossl_b200 import _lib  # noqa: E402
from geossl_b200.NCSN import NCSN_version_03  # noqa: E402
from geossl_b200.data import synthetic_batch  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
b = synthetic_batch(256, 30, seed=0).to(dev)
head = NCSN_version_03(128, 10, 0.01, 50, "symmetry", 2.0).to(dev)
h = torch.randn(b.positions.shape[0], 128, device=dev, requires_grad=True)
sei = b.super_edge_index
d = (b.positions[sei[0]] - b.positions[sei[1]]).norm(dim=1, keepdim=True)
buf = torch.zeros(512, dtype=torch.int64, device=dev)
NAMES = ["start", "scalars", "emb", "feat", "mma1", "e1", "mma2", "e2+loss", "dz2", "mma3", "e3", "mma4", "e4", "end"]


def show(tag):
    t = buf.cpu().view(32, 16)
    print(tag + " tile | phase durations in cycles: " + " ".join(f"{n:>8s}" for n in NAMES[1:]) + " |    total")
    for i in range(8):
        row = [int(v) for v in t[i][:14]]
        if row[0] == 0:
            break
        last = row[0]
        cells = []
        for k in range(1, 14):
            if row[k] > 0:
                cells.append(f"{row[k] - last:8d}")
                last = row[k]
            else:
                cells.append(f"{'-':>8s}")
        print(f"{tag} {i:4d} |                            " + " ".join(cells) + f" | {last - row[0]:8d}")


for it in range(4):
    if it == 3:
        buf.zero_()
        _lib.check(lib.geossl_debug_set_trace_head(ctypes.c_void_p(buf.data_ptr())))
    loss = head(b, h, d)
    if it == 3:
        torch.cuda.synchronize()
        show("fwd")
        buf.zero_()
    loss.backward()
torch.cuda.synchronize()
show("bwd")
_lib.check(lib.geossl_debug_set_trace_head(None))
