import torch
try:
    a=torch.cuda.Event(enable_timing=True, external=True); b=torch.cuda.Event(enable_timing=True, external=True)
    x=torch.randn(4096,4096,device='cuda')
    s=torch.cuda.Stream()
    with torch.cuda.stream(s):
        y=x@x
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        a.record(); y=x@x; b.record()
    for _ in range(3):
        g.replay(); torch.cuda.synchronize(); print("external events in graph:", a.elapsed_time(b))
except Exception as e:
    print("external event probe failed:", repr(e))
