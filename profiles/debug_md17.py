import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from geossl_b200.Geom3D.models import SchNet
from geossl_b200.finetune import GraphedMD17Step, md17_train_step
from geossl_b200.pretrain import default_args
dev = "cuda:0"
torch.manual_seed(0)
m = SchNet(hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=10.0, node_class=9).to(dev)
lin = torch.nn.Linear(128, 1).to(dev)
crit = torch.nn.L1Loss()
opt = torch.optim.Adam(list(m.parameters()) + list(lin.parameters()), lr=5e-4, fused=True, capturable=True)
pool = [bench._ft_batch("md17", i).to(dev) for i in range(4)]
targs = default_args("schnet")
for i in range(3):
    md17_train_step(targs, pool[i], m, lin, crit, opt)
torch.cuda.synchronize()
step = GraphedMD17Step(targs, pool[0], m, lin, crit, opt)
for i in range(3):
    step(pool[i]); torch.cuda.synchronize()
print("graphs", len(step.graphs), list(step.graphs))
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
print("full step ms", t(lambda i: step(pool[i % 4])))
print("structure ms", t(lambda i: step._structure(pool[i % 4])))
g, static, sg, loss = next(iter(step.graphs.values()))
print("replay only ms", t(lambda i: g.replay()))
print("eager ms", t(lambda i: md17_train_step(targs, pool[i % 4], m, lin, crit, opt)))
