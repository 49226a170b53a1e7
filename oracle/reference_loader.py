"""Imports the UNMODIFIED reference modules under the import shims in oracle/shims.  TEST INFRASTRUCTURE.

Where the files come from: /root/reference (the build container) or, on the GPU box, the git-ignored copy that
``oracle/make_ref.py`` placed in ``oracle/_ref`` (same bytes).  Used by tests/golden/make_golden.py, by the optional
cross-checks in tests/test_oracle_golden.py and by the CPU legs of bench.py (``cpu_baseline`` / ``--impl reference``).

``do_DDM`` / ``perturb`` themselves cannot be imported: examples/pretrain_GeoSSL.py parses ``sys.argv`` at import
(config.py:214), imports a missing ``AutoEncoder`` (:17) and PyG's DataLoader (:12).  ``ddm_step`` below therefore
drives the imported reference MODULES with do_DDM's call sequence (pretrain_GeoSSL.py:68-74,179-212).
"""
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.environ.get("GEOSSL_REFERENCE_ROOT", "/root/reference"), os.path.join(_HERE, "_ref"))


def reference_root():
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, "Geom3D", "models", "schnet.py")) and \
                os.path.isfile(os.path.join(root, "examples", "NCSN.py")):
            return root
    return None


REFERENCE_ROOT = reference_root()


def available():
    return reference_root() is not None


def load():
    """Returns (SchNet, PaiNN, NCSN_version_03) classes of the reference."""
    root = reference_root()
    if root is None:
        raise RuntimeError(f"reference modules not found (looked in {_CANDIDATES}); run oracle/make_ref.py in the build container")
    repo = os.path.dirname(_HERE)
    for p in (repo, os.path.join(_HERE, "shims"), root, os.path.join(root, "examples")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the product package also ships a drop-in `Geom3D`; make sure the reference's one wins here
    for name in [m for m in sys.modules if m == "Geom3D" or m.startswith("Geom3D.")]:
        del sys.modules[name]
    sys.path.remove(root)
    sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from Geom3D.models import SchNet, PaiNN          # noqa: E402
        from NCSN import NCSN_version_03                 # noqa: E402
    return SchNet, PaiNN, NCSN_version_03


def ddm_step(model_3d, model, heads, batch, mu=0.0, sigma=0.3):
    """One do_DDM evaluation on the reference modules, every random draw made by the reference's own torch calls
    in the reference's order: perturb's CPU ``torch.normal`` (pretrain_GeoSSL.py:72), then ``randint`` / ``randn_like``
    inside each NCSN head (NCSN.py:190,194).  ``batch`` duck-types BatchAtomTuple.  Returns the loss (:210)."""
    import torch
    x_01 = batch.x[:, 0]                                                                  # :180
    positions_01 = batch.positions
    positions_02 = positions_01 + torch.normal(mu, sigma, size=positions_01.size())       # :68-74
    if model_3d == "schnet":                                                              # :186-188
        _, repr_01 = model(x_01, positions_01, batch.batch, return_latent=True)
        _, repr_02 = model(x_01, positions_02, batch.batch, return_latent=True)
    else:                                                                                 # :189-191
        _, repr_01 = model(x_01, positions_01, batch.radius_edge_index, batch.batch, return_latent=True)
        _, repr_02 = model(x_01, positions_02, batch.radius_edge_index, batch.batch, return_latent=True)
    sei = batch.super_edge_index                                                          # :197-205
    u1, v1 = torch.index_select(positions_01, 0, sei[0]), torch.index_select(positions_01, 0, sei[1])
    u2, v2 = torch.index_select(positions_02, 0, sei[0]), torch.index_select(positions_02, 0, sei[1])
    distance_01 = torch.sqrt(torch.sum((u1 - v1) ** 2, dim=1)).unsqueeze(1)
    distance_02 = torch.sqrt(torch.sum((u2 - v2) ** 2, dim=1)).unsqueeze(1)
    loss_01 = heads[0](batch, repr_01, distance_02)                                       # :207
    loss_02 = heads[1](batch, repr_02, distance_01)                                       # :208
    return (loss_01 + loss_02) / 2                                                        # :210


def md17_step(model, graph_pred_linear, batch, y, force, energy_coeff=0.05, force_coeff=0.95):
    """The loss of finetune_md17.py::train (:30-51) on the reference SchNet: energy from the readout, force by
    ``autograd.grad(..., create_graph=True)``, weighted L1 (coefficients of submit_finetune_md17_schnet.sh)."""
    import torch
    positions = batch.positions.clone().requires_grad_()                                  # :32-33
    x = batch.x[:, 0] if batch.x.dim() == 2 else batch.x
    molecule_3D_repr = model(x, positions, batch.batch)                                   # :37
    pred_energy = graph_pred_linear(molecule_3D_repr).squeeze(1)                          # :42
    pred_force = -torch.autograd.grad(outputs=pred_energy, inputs=positions, grad_outputs=torch.ones_like(pred_energy),
                                      create_graph=True, retain_graph=True)[0]            # :46
    crit = torch.nn.L1Loss()
    return energy_coeff * crit(pred_energy, y) + force_coeff * crit(pred_force, force)    # :51


def lba_step(model, graph_pred_linear, batch, y):
    """The loss of finetune_lba.py::train (:33-47): SchNet readout -> Linear -> MSE (criterion, :244)."""
    import torch
    x = batch.x[:, 0] if batch.x.dim() == 2 else batch.x
    molecule_3D_repr = model(x, batch.positions, batch.batch)                             # :37
    pred = graph_pred_linear(molecule_3D_repr).squeeze()                                  # :42
    return torch.nn.MSELoss()(pred, y)                                                    # :47
