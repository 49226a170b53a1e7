"""Per-key parity report of the CUDA path against the reference-generated fixtures at the BENCHMARKED shapes
(tests/golden/ddm_schnet_cfg1.npz = 32 x 30 atoms, ddm_schnet_cfg2.npz = 256 x 30 atoms; both written by the unmodified
reference modules, tests/golden/make_golden.py).  Prints max-norm relative errors of the loss and of every parameter
gradient, per kernel mode; the bounds asserted in tests/test_gpu_ddm.py come from this table.

    python profiles/parity_report.py > profiles/rNN_parity_report.txt
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]

from _build import grads_of, head_from, schnet_from  # noqa: E402
from _golden import Golden, rel_err  # noqa: E402
from geossl_b200 import ops  # noqa: E402
from geossl_b200.data import AtomTupleBatch  # noqa: E402
from geossl_b200.pretrain import default_args, do_DDM  # noqa: E402

DEV = "cuda:0"


def run(name, mode, stack):
    ops.FILTER_MODE = mode
    g = Golden(name)
    c, i = g.cfg, g["in"]
    model = schnet_from(g, DEV)
    heads = (head_from(g, "sd1", DEV), head_from(g, "sd2", DEV))
    batch = AtomTupleBatch(i["x"].to(DEV), i["pos"].to(DEV), i["batch"].to(DEV), i["super_edge_index"].to(DEV), None,
                           n_graphs=int(i["batch"][-1]) + 1)
    draws = ((i["noise_level_1"].to(DEV), i["distance_noise_1"].to(DEV)),
             (i["noise_level_2"].to(DEV), i["distance_noise_2"].to(DEV)))
    loss, _ = do_DDM(default_args("schnet"), batch, model, None, 0.0, c["sigma"], heads=heads, draws=draws,
                     positions_02=(i["pos"] + i["pos_noise"]).to(DEV), stack_views=stack)
    loss.backward()
    torch.cuda.synchronize()
    rows = [("loss", rel_err(loss, g["out"]["loss"]))]
    for mod, grp in ((model, "grad"), (heads[0], "grad1"), (heads[1], "grad2")):
        got = grads_of(mod)
        for k, ref in g[grp].items():
            rows.append((f"{grp}/{k}", rel_err(got[k], ref)))
    return rows


if __name__ == "__main__":
    for name in ("ddm_schnet_full4", "ddm_schnet_cfg1", "ddm_schnet_cfg2"):
        for mode in ("simt", "tc_fp16", "tc_bf16"):
            for stack in (True, False):
                rows = run(name, mode, stack)
                worst = max(rows[1:], key=lambda r: r[1])
                enc = max((r for r in rows[1:] if r[0].startswith("grad/")), key=lambda r: r[1])
                head = max((r for r in rows[1:] if not r[0].startswith("grad/")), key=lambda r: r[1])
                print(f"{name:18s} {mode:8s} stack={int(stack)}  loss {rows[0][1]:.2e}  worst encoder grad {enc[1]:.2e} ({enc[0]})  "
                      f"worst head grad {head[1]:.2e} ({head[0]})")
                if "-v" in sys.argv:
                    for k, e in rows:
                        print(f"    {k:60s} {e:.3e}")
