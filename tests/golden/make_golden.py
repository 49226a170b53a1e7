"""Generates the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference modules
(/root/reference/Geom3D/models/{schnet,painn}.py, /root/reference/examples/NCSN.py) on CPU under the
import shims in oracle/shims.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests / golden vectors of its own, so these are the pins.
``do_DDM`` / ``perturb`` cannot be imported (pretrain_GeoSSL.py parses argv and imports a missing
``AutoEncoder`` at import time), so this script drives the reference *modules* with the same call
sequence as pretrain_GeoSSL.py:179-212 and finetune_md17.py:32-54.
The radius graph comes from oracle/radius.py (torch_cluster is absent: parity unpinned for it).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import reference_loader  # noqa: E402
from geossl_b200.data import synthetic_batch  # noqa: E402

SchNet, PaiNN, NCSN_version_03 = reference_loader.load()
from oracle.radius import radius_graph  # noqa: E402

torch.set_num_threads(4)


def flat(prefix, d):
    # the aliased ``conv.nn.*`` keys (schnet.py:141-148,175) duplicate ``mlp.*`` bit for bit: not stored
    return {f"{prefix}/{k}": (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
            for k, v in d.items() if ".conv.nn." not in k}


def save(name, cfg, **groups):
    out = {"cfg": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)}
    for g, d in groups.items():
        out.update(flat(g, d))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(out)} arrays")


def grads_of(module):
    return {k: p.grad for k, p in module.named_parameters() if p.grad is not None}


def schnet_case(name, *, seed, hidden, filters, gaussians, layers, cutoff, readout, num_graphs, atoms,
                atoms_max=None, density=0.05):
    torch.manual_seed(seed)
    model = SchNet(hidden_channels=hidden, num_filters=filters, num_interactions=layers,
                   num_gaussians=gaussians, cutoff=cutoff, node_class=9, readout=readout)
    b = synthetic_batch(num_graphs, atoms, atoms_max, seed=seed + 1, density=density)
    z = b.x[:, 0].contiguous()
    out, h = model(z, b.positions, b.batch, return_latent=True)
    g = torch.Generator().manual_seed(seed + 2)
    w_h = torch.randn(h.shape, generator=g)
    w_o = torch.randn(out.shape, generator=g)
    loss = (h * w_h).sum() + (out * w_o).sum()
    loss.backward()
    ei = radius_graph(b.positions, cutoff, b.batch)
    cfg = dict(kind="schnet", hidden=hidden, filters=filters, gaussians=gaussians, layers=layers,
               cutoff=cutoff, readout=readout, seed=seed)
    save(name, cfg,
         **{"in": dict(z=z, pos=b.positions, batch=b.batch, w_h=w_h, w_o=w_o),
            "sd": dict(model.state_dict()),
            "out": dict(out=out, h=h, loss=loss, edge_index=ei),
            "grad": grads_of(model)})


def painn_case(name, *, seed, feat, layers, rbf, cutoff, readout, num_graphs, atoms, atoms_max=None):
    torch.manual_seed(seed)
    model = PaiNN(n_atom_basis=feat, n_interactions=layers, n_rbf=rbf, cutoff=cutoff, max_z=9, n_out=1,
                  readout=readout)
    b = synthetic_batch(num_graphs, atoms, atoms_max, seed=seed + 1)
    rei = radius_graph(b.positions, cutoff, b.batch)
    h, q = model(b.x, b.positions, rei, b.batch, return_latent=True)
    g = torch.Generator().manual_seed(seed + 2)
    w_q = torch.randn(q.shape, generator=g)
    w_h = torch.randn(h.shape, generator=g)
    loss = (q * w_q).sum() + (h * w_h).sum()
    loss.backward()
    cfg = dict(kind="painn", feat=feat, layers=layers, rbf=rbf, cutoff=cutoff, readout=readout, seed=seed)
    save(name, cfg,
         **{"in": dict(x=b.x, pos=b.positions, batch=b.batch, radius_edge_index=rei, w_q=w_q, w_h=w_h),
            "sd": dict(model.state_dict()),
            "out": dict(h=h, q=q, loss=loss),
            "grad": grads_of(model)})


def ncsn_case(name, *, seed, emb, levels, anneal_power, num_graphs, atoms, atoms_max=None,
              option="combination"):
    torch.manual_seed(seed)
    head = NCSN_version_03(emb, sigma_begin=10, sigma_end=0.01, num_noise_level=levels,
                           noise_type="symmetry", anneal_power=anneal_power)
    b = synthetic_batch(num_graphs, atoms, atoms_max, seed=seed + 1, option=option)
    g = torch.Generator().manual_seed(seed + 2)
    node_feature = torch.randn((b.positions.size(0), emb), generator=g).requires_grad_()
    u, v = b.positions[b.super_edge_index[0]], b.positions[b.super_edge_index[1]]
    distance = torch.sqrt(torch.sum((u - v) ** 2, dim=1)).unsqueeze(1)
    torch.manual_seed(seed + 3)
    loss = head(b, node_feature, distance)
    loss.backward()
    torch.manual_seed(seed + 3)               # replay the two draws of NCSN.py:190,194
    noise_level = torch.randint(0, levels, (b.num_graphs,))
    distance_noise = torch.randn_like(distance)
    cfg = dict(kind="ncsn", emb=emb, levels=levels, anneal_power=anneal_power, seed=seed, option=option)
    save(name, cfg,
         **{"in": dict(batch=b.batch, super_edge_index=b.super_edge_index, node_feature=node_feature,
                       distance=distance, noise_level=noise_level, distance_noise=distance_noise),
            "sd": dict(head.state_dict()),
            "out": dict(loss=loss),
            "grad": dict(node_feature=node_feature.grad, **grads_of(head))})


def ddm_case(name, *, seed, model_3d, emb, num_graphs, atoms, atoms_max=None, levels=50, anneal_power=2.0,
             sigma=0.3, store_reprs=True, **enc):
    """The DDM step of pretrain_GeoSSL.py:179-212 driven on the reference modules.  ``store_reprs=False`` drops the two
    (N,H) representations from the file (the bench-sized fixture would otherwise be 8 MB larger); loss, both head losses
    and every parameter gradient are always stored."""
    torch.manual_seed(seed)
    if model_3d == "schnet":
        model = SchNet(hidden_channels=emb, num_filters=enc["filters"], num_interactions=enc["layers"],
                       num_gaussians=enc["gaussians"], cutoff=enc["cutoff"], node_class=9, readout="mean")
    else:
        model = PaiNN(n_atom_basis=emb, n_interactions=enc["layers"], n_rbf=enc["rbf"], cutoff=enc["cutoff"],
                      max_z=9, n_out=1, readout="add")
    heads = [NCSN_version_03(emb, sigma_begin=10, sigma_end=0.01, num_noise_level=levels,
                             noise_type="symmetry", anneal_power=anneal_power) for _ in range(2)]
    b = synthetic_batch(num_graphs, atoms, atoms_max, seed=seed + 1)
    if model_3d == "painn":
        b.radius_edge_index = radius_graph(b.positions, enc["cutoff"], b.batch)
    torch.manual_seed(seed + 3)
    pos_noise = torch.normal(0.0, sigma, size=b.positions.size())          # perturb(), :72
    x_01 = b.x[:, 0]
    positions_01 = b.positions
    positions_02 = positions_01 + pos_noise
    if model_3d == "schnet":
        _, repr_01 = model(x_01, positions_01, b.batch, return_latent=True)
        _, repr_02 = model(x_01, positions_02, b.batch, return_latent=True)
    else:
        _, repr_01 = model(x_01, positions_01, b.radius_edge_index, b.batch, return_latent=True)
        _, repr_02 = model(x_01, positions_02, b.radius_edge_index, b.batch, return_latent=True)
    sei = b.super_edge_index
    d01 = torch.sqrt(torch.sum((positions_01[sei[0]] - positions_01[sei[1]]) ** 2, dim=1)).unsqueeze(1)
    d02 = torch.sqrt(torch.sum((positions_02[sei[0]] - positions_02[sei[1]]) ** 2, dim=1)).unsqueeze(1)
    state = torch.get_rng_state()
    loss_01 = heads[0](b, repr_01, d02)
    loss_02 = heads[1](b, repr_02, d01)
    loss = (loss_01 + loss_02) / 2
    loss.backward()
    torch.set_rng_state(state)                                             # replay the four draws
    nl1 = torch.randint(0, levels, (b.num_graphs,)); eps1 = torch.randn_like(d02)
    nl2 = torch.randint(0, levels, (b.num_graphs,)); eps2 = torch.randn_like(d01)
    cfg = dict(kind="ddm", model_3d=model_3d, emb=emb, levels=levels, anneal_power=anneal_power, sigma=sigma,
               seed=seed, **enc)
    ins = dict(x=b.x, pos=b.positions, batch=b.batch, super_edge_index=sei, pos_noise=pos_noise,
               noise_level_1=nl1, distance_noise_1=eps1, noise_level_2=nl2, distance_noise_2=eps2)
    if model_3d == "painn":
        ins["radius_edge_index"] = b.radius_edge_index
    save(name, cfg,
         **{"in": ins, "sd": dict(model.state_dict()), "sd1": dict(heads[0].state_dict()),
            "sd2": dict(heads[1].state_dict()),
            "out": dict(loss=loss, loss_01=loss_01, loss_02=loss_02,
                        **(dict(repr_01=repr_01, repr_02=repr_02) if store_reprs else {})),
            "grad": grads_of(model), "grad1": grads_of(heads[0]), "grad2": grads_of(heads[1])})


def md17_case(name, *, seed, hidden, filters, gaussians, layers, cutoff, num_graphs, atoms):
    """finetune_md17.py:32-54: energy + autograd force, loss backward (double backward)."""
    torch.manual_seed(seed)
    model = SchNet(hidden_channels=hidden, num_filters=filters, num_interactions=layers,
                   num_gaussians=gaussians, cutoff=cutoff, node_class=9, readout="mean")
    lin = torch.nn.Linear(hidden, 1)
    b = synthetic_batch(num_graphs, atoms, seed=seed + 1, density=0.08)
    z = b.x[:, 0].contiguous()
    g = torch.Generator().manual_seed(seed + 2)
    y = torch.randn(num_graphs, generator=g)
    f_t = torch.randn(b.positions.shape, generator=g)
    pos = b.positions.clone().requires_grad_()
    rep = model(z, pos, b.batch)
    energy = lin(rep).squeeze(1)
    force = -torch.autograd.grad(energy, pos, grad_outputs=torch.ones_like(energy), create_graph=True,
                                 retain_graph=True)[0]
    loss = 0.05 * torch.nn.functional.l1_loss(energy, y) + 0.95 * torch.nn.functional.l1_loss(force, f_t)
    loss.backward()
    cfg = dict(kind="md17", hidden=hidden, filters=filters, gaussians=gaussians, layers=layers, cutoff=cutoff,
               readout="mean", seed=seed)
    save(name, cfg,
         **{"in": dict(z=z, pos=b.positions, batch=b.batch, y=y, force_target=f_t),
            "sd": dict(model.state_dict()), "sdlin": dict(lin.state_dict()),
            "out": dict(energy=energy, force=force, loss=loss),
            "grad": grads_of(model), "gradlin": grads_of(lin)})


def ssl_case(name, *, seed, hidden, filters, gaussians, layers, cutoff, num_graphs, atoms, atoms_max=None,
             T=0.1, num_neg=1, sigma=0.3):
    """The sibling objectives on the reference SchNet: do_InfoNCE (pretrain_GeoSSL.py:141-176), do_EBM_NCE (:103-138,
    criterion = BCEWithLogitsLoss :344) and the DistancePredictor step (pretrain_DistancePrediction.py:15-26,64-79).
    Each objective's loss is back-propagated separately into the same encoder."""
    torch.manual_seed(seed)
    model = SchNet(hidden_channels=hidden, num_filters=filters, num_interactions=layers,
                   num_gaussians=gaussians, cutoff=cutoff, node_class=9, readout="mean")
    predictor = torch.nn.Linear(hidden * 2, 1)                              # DistancePredictor.predictor, :18
    b = synthetic_batch(num_graphs, atoms, atoms_max, seed=seed + 1)
    torch.manual_seed(seed + 3)
    pos_noise = torch.normal(0.0, sigma, size=b.positions.size())          # perturb(), :72
    x_01 = b.x[:, 0]
    positions_02 = b.positions + pos_noise
    out, grads = {}, {}

    def reprs():
        return model(x_01, b.positions, b.batch), model(x_01, positions_02, b.batch)

    # ---- InfoNCE, :159-176
    r1, r2 = reprs()
    CE = torch.nn.CrossEntropyLoss()

    def cal_loss(X, Y):
        B = X.size()[0]
        logits = torch.div(torch.mm(X, Y.transpose(1, 0)), T)
        labels = torch.arange(B).long()
        return CE(logits, labels), logits.argmax(dim=1).eq(labels).sum().item() * 1. / B
    l01, a01 = cal_loss(r1, r2)
    l02, a02 = cal_loss(r2, r1)
    loss = (l01 + l02) / 2
    model.zero_grad(); loss.backward()
    out.update(repr_01=r1, repr_02=r2, infonce_loss=loss, infonce_acc=torch.tensor((a01 + a02) / 2))
    grads.update({"infonce/" + k: v.clone() for k, v in grads_of(model).items()})

    # ---- EBM-NCE, :122-138
    r1, r2 = reprs()
    B = len(r1)
    cyc = lambda num, shift: torch.cat([torch.arange(shift, num), torch.arange(shift)])   # == util.py:19-22
    neg_01 = r1.repeat((num_neg, 1))
    neg_02 = torch.cat([r2[cyc(B, i + 1)] for i in range(num_neg)], dim=0)
    pred_pos = torch.sum(r1 * r2, dim=1)
    pred_neg = torch.sum(neg_01 * neg_02, dim=1)
    BCE = torch.nn.BCEWithLogitsLoss()
    loss_pos = BCE(pred_pos.double(), torch.ones(B).double())
    loss_neg = BCE(pred_neg.double(), torch.zeros(B * num_neg).double())
    loss = (loss_pos + num_neg * loss_neg) / (1 + num_neg)
    acc = (torch.sum(pred_pos > 0).float() + torch.sum(pred_neg < 0).float()) / (len(pred_pos) + len(pred_neg))
    model.zero_grad(); loss.backward()
    out.update(ebm_loss=loss, ebm_acc=acc)
    grads.update({"ebm/" + k: v.clone() for k, v in grads_of(model).items()})

    # ---- distance prediction, pretrain_DistancePrediction.py:64-79
    _, node_repr = model(x_01, b.positions, b.batch, return_latent=True)
    sei = b.super_edge_index
    u_node_repr = torch.index_select(node_repr, dim=0, index=sei[0])
    v_node_repr = torch.index_select(node_repr, dim=0, index=sei[1])
    u_pos = torch.index_select(b.positions, dim=0, index=sei[0])
    v_pos = torch.index_select(b.positions, dim=0, index=sei[1])
    distance_actual = torch.sqrt(torch.sum((u_pos - v_pos) ** 2, dim=1))
    distance_pred = predictor(torch.cat([u_node_repr, v_node_repr], dim=1)).squeeze()
    loss = torch.nn.L1Loss()(distance_pred, distance_actual)
    model.zero_grad(); predictor.zero_grad(); loss.backward()
    out.update(distance_loss=loss)
    grads.update({"distance/" + k: v.clone() for k, v in grads_of(model).items()})
    grads.update({"distance_predictor/" + k: v.clone() for k, v in grads_of(predictor).items()})

    cfg = dict(kind="ssl", hidden=hidden, filters=filters, gaussians=gaussians, layers=layers, cutoff=cutoff,
               readout="mean", T=T, num_neg=num_neg, sigma=sigma, seed=seed)
    save(name, cfg,
         **{"in": dict(x=b.x, pos=b.positions, batch=b.batch, super_edge_index=sei, pos_noise=pos_noise),
            "sd": dict(model.state_dict()), "sdpred": {"predictor." + k: v for k, v in predictor.state_dict().items()},
            "out": out, "grad": grads})


def masking_case(name, *, seed, mask_ratio, sizes):
    """``Molecule3DDataset.subgraph`` (datasets_3D.py:24-67) of the UNMODIFIED reference file (imported under the
    torch_geometric.utils / .data shims) on synthetic molecules: a random spanning tree + a few ring bonds as the
    (bidirectional) bond graph, bond attributes, and a dataset-time radius graph.  Stores inputs, the numpy seed and the
    masked molecules."""
    import importlib.util
    from torch_geometric.data import Data
    spec = importlib.util.spec_from_file_location("ref_datasets_3D", os.path.join(reference_loader.reference_root(),
                                                                                  "Geom3D", "datasets", "datasets_3D.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ds = ref.Molecule3DDataset.__new__(ref.Molecule3DDataset)
    ds.mask_ratio = mask_ratio
    rng = np.random.default_rng(seed)
    ins, outs = {}, {}
    np.random.seed(seed)
    for m, n in enumerate(sizes):
        parents = [int(rng.integers(0, i)) for i in range(1, n)]
        und = [(i + 1, p) for i, p in enumerate(parents)]
        for _ in range(max(1, n // 6)):                                     # ring closures
            a, b = (int(v) for v in rng.choice(n, 2, replace=False))
            if (a, b) not in und and (b, a) not in und:
                und.append((a, b))
        if m == 1:                                                          # one molecule with a disconnected fragment
            und = [e for e in und if n - 1 not in e and n - 2 not in e] + [(n - 1, n - 2)]
        ei = torch.tensor([[a, b] for a, b in und] + [[b, a] for a, b in und], dtype=torch.long).t().contiguous()
        ei = ei[:, torch.argsort(ei[0] * n + ei[1])]
        ea = torch.from_numpy(rng.integers(0, 4, size=(ei.size(1), 2)))
        x = torch.from_numpy(np.stack([rng.integers(0, 9, n), np.zeros(n, np.int64)], 1))
        pos = torch.from_numpy(rng.random((n, 3)).astype(np.float32) * 6.0)
        rei = radius_graph(pos, 3.0, torch.zeros(n, dtype=torch.long))
        ins.update({f"x{m}": x, f"positions{m}": pos, f"edge_index{m}": ei, f"edge_attr{m}": ea, f"radius_edge_index{m}": rei})
        d = ds.subgraph(Data(x=x.clone(), positions=pos.clone(), edge_index=ei.clone(), edge_attr=ea.clone(),
                             radius_edge_index=rei.clone()))
        outs.update({f"x{m}": d.x, f"positions{m}": d.positions, f"edge_index{m}": d.edge_index, f"edge_attr{m}": d.edge_attr,
                     f"radius_edge_index{m}": d.radius_edge_index})
    save(name, dict(kind="masking", seed=seed, mask_ratio=mask_ratio, n_mol=len(sizes)), **{"in": ins, "out": outs})


if __name__ == "__main__":
    only = set(sys.argv[1:])            # optional: names of the fixtures to (re)generate

    def _wrap(fn):
        return lambda name, **kw: fn(name, **kw) if (not only or name in only) else None
    schnet_case, painn_case, ncsn_case, ddm_case, md17_case, ssl_case, masking_case = map(
        _wrap, (schnet_case, painn_case, ncsn_case, ddm_case, md17_case, ssl_case, masking_case))
    schnet_case("schnet_small", seed=11, hidden=32, filters=32, gaussians=20, layers=2, cutoff=10.0,
                readout="mean", num_graphs=5, atoms=4, atoms_max=12)
    schnet_case("schnet_trunc", seed=12, hidden=32, filters=64, gaussians=51, layers=2, cutoff=10.0,
                readout="add", num_graphs=3, atoms=36, atoms_max=60, density=0.08)
    # (full-size SchNet, H=F=128 / G=50 / L=6, is pinned by ddm_schnet_full4 / _cfg1 / _cfg2 below)
    painn_case("painn_small", seed=21, feat=32, layers=2, rbf=20, cutoff=5.0, readout="add",
               num_graphs=5, atoms=4, atoms_max=14)
    painn_case("painn_full", seed=22, feat=128, layers=3, rbf=20, cutoff=5.0, readout="add",
               num_graphs=4, atoms=30)
    ncsn_case("ncsn_h128", seed=31, emb=128, levels=50, anneal_power=2.0, num_graphs=6, atoms=3, atoms_max=16)
    ncsn_case("ncsn_perm", seed=32, emb=32, levels=30, anneal_power=0.05, num_graphs=4, atoms=2, atoms_max=9,
              option="permutation")
    ddm_case("ddm_schnet_small", seed=41, model_3d="schnet", emb=32, num_graphs=6, atoms=5, atoms_max=14,
             filters=32, gaussians=20, layers=2, cutoff=10.0)
    ddm_case("ddm_schnet_full4", seed=42, model_3d="schnet", emb=128, num_graphs=4, atoms=30,
             filters=128, gaussians=50, layers=6, cutoff=10.0)
    # BASELINE.json configs[0] and configs[1] at their real batch sizes (32 x 30 and 256 x 30 atoms, full model)
    ddm_case("ddm_schnet_cfg1", seed=44, model_3d="schnet", emb=128, num_graphs=32, atoms=30,
             filters=128, gaussians=50, layers=6, cutoff=10.0)
    ddm_case("ddm_schnet_cfg2", seed=45, model_3d="schnet", emb=128, num_graphs=256, atoms=30, store_reprs=False,
             filters=128, gaussians=50, layers=6, cutoff=10.0)
    ddm_case("ddm_painn_small", seed=43, model_3d="painn", emb=32, num_graphs=6, atoms=5, atoms_max=14,
             layers=2, rbf=20, cutoff=5.0)
    md17_case("md17_small", seed=51, hidden=32, filters=32, gaussians=20, layers=2, cutoff=10.0,
              num_graphs=3, atoms=9)
    ssl_case("ssl_schnet_small", seed=61, hidden=32, filters=32, gaussians=20, layers=2, cutoff=10.0,
             num_graphs=7, atoms=4, atoms_max=12)
    masking_case("masking_small", seed=71, mask_ratio=0.3, sizes=[12, 17, 9, 30, 5, 23])
