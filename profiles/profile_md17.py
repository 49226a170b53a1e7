"""One eager MD17-shaped force fine-tune step (configs[3]) between cudaProfilerStart/Stop, for the ncu launch list.
Use under ncu with `--profile-from-start off`.  Never a bench number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from geossl_b200.Geom3D.models import SchNet  # noqa: E402
from geossl_b200.finetune import md17_train_step  # noqa: E402
from geossl_b200.pretrain import default_args  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(42)
f = bench.FT["md17"]
model = SchNet(hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=f["cutoff"], node_class=9).to(dev)
lin = torch.nn.Linear(128, 1).to(dev)
params = list(model.parameters()) + list(lin.parameters())
opt = torch.optim.Adam(params, lr=5e-4, fused=True)
pool = [bench._ft_batch("md17", i).to(dev) for i in range(3)]
crit = torch.nn.L1Loss()
targs = default_args("schnet")
for i in range(3):
    md17_train_step(targs, pool[i % 3], model, lin, crit, opt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
md17_train_step(targs, pool[0], model, lin, crit, opt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled 1 step")
