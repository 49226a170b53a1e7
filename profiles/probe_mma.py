"""tcgen05.mma issue/throughput probe (selftest mode 2): cycles per M128 x N x K16 kind::f16 SS-mode MMA."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geossl_b200 import _lib
lib = _lib.load()
NAMES = {2: "SS  K-major A,B from smem", 4: "TS  A in TMEM, B K-major smem", 5: "SS  A,B MN-major from smem",
         6: "SS  A MN-major, B K-major", 7: "SS  A K-major, B MN-major"}
for mode, N in ((2, 128), (2, 64), (4, 128), (4, 64), (5, 128), (5, 64), (6, 128), (6, 64), (7, 128), (7, 64)):
    a, b = torch.randn(128, 64, device="cuda"), torch.randn(128, 64, device="cuda")
    d = torch.zeros(128, N, device="cuda")
    for _ in range(2):
        _lib.check(lib.geossl_tc_selftest(mode, 1, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), 64, N,
                                          ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
    t = d.view(-1)[:4].view(torch.int64).cpu()
    print(f"{NAMES[mode]:32s} N={N}: issue {int(t[0]) / 240:.1f} cyc/MMA, complete {int(t[1]) / 240:.1f} cyc/MMA (240 MMAs, idle SM)")
