#!/bin/bash
# compute-sanitizer pass over a small end-to-end slice of every kernel family (sanitizer-clean subset).
#   gpurun --timeout 1500 -- 'bash profiles/scripts/sanitize.sh TAG'   ->  gpurun_out/sanitizer_TAG.txt  (copied to profiles/)
TAG=${1:-x}
OUT=gpurun_out/sanitizer_$TAG.txt
mkdir -p gpurun_out
: > $OUT
run() {   # tool, label, python snippet
    echo "===== compute-sanitizer --tool $1 : $2" >> $OUT
    timeout 600 compute-sanitizer --tool $1 --print-limit 20 --error-exitcode 9 python -c "$3" > gpurun_out/san_tmp.txt 2>&1
    echo "exit code $?" >> $OUT
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race reported|smoke:|mini-step" gpurun_out/san_tmp.txt | sed -E "s/\(.*\)\+0x[0-9a-f]+//" | sort | uniq -c | sort -rn | head -30 >> $OUT
}
SMOKE='import __graft_entry__ as g; g.smoke()'
MINI='
import torch
from geossl_b200.Geom3D.models import PaiNN, SchNet
from geossl_b200.NCSN import NCSN_version_03
from geossl_b200.data import synthetic_batch, pad_batch
from geossl_b200 import ops
from geossl_b200.pretrain import default_args, do_DDM
dev = "cuda:0"
torch.manual_seed(0)
b = synthetic_batch(6, 8, 40, seed=1).to(dev)
heads = [NCSN_version_03(128, 10, 0.01, 50, "symmetry", 2.0).to(dev) for _ in range(2)]
m = SchNet(node_class=9, num_interactions=2).to(dev)
loss, _ = do_DDM(default_args("schnet"), pad_batch(b, b.positions.size(0) + 9, b.super_edge_index.size(1) + 100), m, None, heads=heads, device_noise=True)
loss.backward()
p = PaiNN(n_atom_basis=128, n_interactions=2, n_rbf=20, cutoff=5.0, max_z=9, n_out=1, readout="add").to(dev)
b.radius_edge_index = ops.radius_csr(b.positions, b.batch, 5.0, num_graphs=6, transpose=False).edge_index
b.extras["rei_sorted"] = True
loss2, _ = do_DDM(default_args("painn"), b, p, None, heads=heads, device_noise=True)
loss2.backward()
g = ops.radius_csr(b.positions, b.batch, 10.0, num_graphs=6, cell_list=True)
# pair-centric cfconv kernel (graphs of <= 32 atoms, incl. one above the bound -> in-kernel global path) and the MD17 double-backward
# step on the tensor-core products closed under differentiation + per-pair filter rows
from geossl_b200.finetune import md17_train_step
s = synthetic_batch(5, 9, 30, seed=3)
s.extras["max_graph_atoms"] = 24
s = s.to(dev)
loss3, _ = do_DDM(default_args("schnet"), s, m, None, heads=heads, device_noise=True)
loss3.backward()
f = synthetic_batch(4, 21, seed=5, with_pairs=False).to(dev)
f.extras["y"], f.extras["force"] = torch.randn(4, device=dev), torch.randn_like(f.positions)
lin = torch.nn.Linear(128, 1).to(dev)
opt = torch.optim.Adam(list(m.parameters()) + list(lin.parameters()), lr=1e-4)
loss4 = md17_train_step(default_args("schnet"), f, m, lin, torch.nn.L1Loss(), opt)
torch.cuda.synchronize()
print("mini-step: schnet (padded) loss", float(loss), "painn loss", float(loss2), "edges", g.num_edges, "pairs-kernel loss", float(loss3),
      "md17 loss", float(loss4))
'
run memcheck "smoke() = SchNet-DDM forward+backward vs oracle" "$SMOKE"
run memcheck "padded SchNet-DDM step, PaiNN-DDM step (tensor-core Dense blocks), cell list, pair-centric cfconv, MD17 double-backward step" "$MINI"
run racecheck "smoke()" "$SMOKE"
run racecheck "padded SchNet-DDM step, PaiNN-DDM step, cell list, pair-centric cfconv, MD17 double-backward step" "$MINI"
run initcheck "smoke()" "$SMOKE"
cat $OUT
