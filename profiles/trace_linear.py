"""Clock trace of CTA 0 of linear_tc_kernel + event-timed kernel duration (debug aid, not a bench)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geossl_b200 import _lib, ops

dev = "cuda:0"
lib = _lib.load()
n = 15360
x = torch.randn(n, 128, device=dev)
lin = torch.nn.Linear(128, 128).to(dev)
res = torch.randn(n, 128, device=dev)
buf = torch.zeros(16, dtype=torch.int64, device=dev)
NAMES = ["entry", "setup", "staged", "synced", "w_landed", "mma_done", "stored", "exit"]
for pre_ssp, r in ((False, None), (True, res)):
    for _ in range(3):
        ops.linear(x, lin, pre_ssp=pre_ssp, residual=r)
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    image = ops._pack_weight(lin.weight, False, False)
    y = torch.empty_like(x)
    call = lambda: lib.geossl_linear_tc(ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(image.data_ptr()), ctypes.c_void_p(lin.bias.data_ptr()),
                                        1 if pre_ssp else 0, None, None if r is None else ctypes.c_void_p(r.data_ptr()),
                                        ctypes.c_void_p(y.data_ptr()), 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            call()
    g.replay(); torch.cuda.synchronize()
    a.record(); g.replay(); c.record(); torch.cuda.synchronize()
    print(f"pre_ssp={pre_ssp} residual={r is not None}: {a.elapsed_time(c) * 1e3 / 20:.2f} us per launch (20 back-to-back launches in a graph)")
    buf.zero_()
    _lib.check(lib.geossl_debug_set_trace_linear(ctypes.c_void_p(buf.data_ptr())))
    call(); torch.cuda.synchronize()
    _lib.check(lib.geossl_debug_set_trace_linear(None))
    t = [int(v) for v in buf.cpu()[:8]]
    print("   " + "  ".join(f"{NAMES[i]}+{t[i] - t[0]}" for i in range(8)))
