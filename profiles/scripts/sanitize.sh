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
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race reported|hazard|smoke:|mini-step" gpurun_out/san_tmp.txt | head -40 >> $OUT
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
torch.cuda.synchronize()
print("mini-step: schnet (padded) loss", float(loss), "painn loss", float(loss2), "edges", g.num_edges)
'
run memcheck "smoke() = SchNet-DDM forward+backward vs oracle" "$SMOKE"
run memcheck "padded SchNet-DDM step, PaiNN-DDM step (tensor-core Dense blocks), cell-list neighbour search" "$MINI"
run racecheck "smoke()" "$SMOKE"
run racecheck "padded SchNet-DDM step, PaiNN-DDM step, cell list" "$MINI"
run initcheck "smoke()" "$SMOKE"
cat $OUT
