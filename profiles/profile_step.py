"""Runs a few DDM pretraining steps of the bench workload with cudaProfilerStart/Stop around the last ones.
Use under ncu with `--profile-from-start off` (recipes in profiles/README.md).  Never a bench number."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from geossl_b200.Geom3D.models import SchNet  # noqa: E402
from geossl_b200.NCSN import NCSN_version_03  # noqa: E402
from geossl_b200.data import synthetic_batch  # noqa: E402
from geossl_b200.pretrain import default_args, train_step  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--batch", type=int, default=bench.CFG["batch_per_gpu"])
ap.add_argument("--model", default="schnet", choices=["schnet", "painn"])
args = ap.parse_args()

C = bench.CFG
dev = torch.device("cuda:0")
torch.manual_seed(42)
if args.model == "painn":
    from geossl_b200.Geom3D.models import PaiNN
    model = PaiNN(n_atom_basis=C["hidden"], n_interactions=3, n_rbf=20, cutoff=5.0, max_z=9, n_out=1, readout="add").to(dev)
else:
    model = SchNet(hidden_channels=C["hidden"], num_filters=C["filters"], num_interactions=C["interactions"],
                   num_gaussians=C["num_gaussians"], cutoff=C["cutoff"], node_class=9).to(dev)
heads = [NCSN_version_03(C["hidden"], 10, 0.01, C["sigma_levels"], "symmetry", C["anneal_power"]).to(dev) for _ in range(2)]
opt = torch.optim.Adam([{"params": model.parameters()}] + [{"params": [p for p in h.parameters() if p.requires_grad]} for h in heads],
                       lr=C["lr"], fused=True)
pool = [synthetic_batch(args.batch, C["atoms"], seed=i).to(dev) for i in range(4)]
if args.model == "painn":
    from geossl_b200 import ops
    for b in pool:
        b.radius_edge_index = ops.radius_csr(b.positions, b.batch, 5.0, num_graphs=args.batch, transpose=False).edge_index
        b.extras["rei_sorted"] = True
targs = default_args(args.model)
for i in range(args.warmup):
    train_step(targs, pool[i % 4], model, heads, opt, 0.0, C["pos_sigma"], device_noise=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(args.steps):
    train_step(targs, pool[(i + 1) % 4], model, heads, opt, 0.0, C["pos_sigma"], device_noise=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.steps, "step(s)")
