"""GPU: the sibling objectives that share do_DDM's two-view encoder pass against the fixture
produced by the reference SchNet on CPU (tests/golden/make_golden.py::ssl_case)."""
import pytest
import torch

from _build import grads_of, schnet_from
from _golden import Golden, rel_err
from geossl_b200.data import AtomTupleBatch
from geossl_b200.pretrain import (DistancePredictor, cycle_index, default_args, do_DistancePrediction, do_EBM_NCE,
                                  do_InfoNCE)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_OUT, TOL_GRAD = 1e-5, 1e-4


def _setup():
    g = Golden("ssl_schnet_small")
    i = g["in"]
    batch = AtomTupleBatch(i["x"].to(DEV), i["pos"].to(DEV), i["batch"].to(DEV), i["super_edge_index"].to(DEV),
                           None, n_graphs=int(i["batch"][-1]) + 1)
    return g, batch, (i["pos"] + i["pos_noise"]).to(DEV)


def _check(model, g, prefix, tol):
    got = grads_of(model)
    for k, ref in g["grad"].items():
        if k.startswith(prefix):
            assert rel_err(got[k[len(prefix):]], ref) <= tol, (k, rel_err(got[k[len(prefix):]], ref))


@pytest.mark.parametrize("stack", [True, False])
def test_info_nce_vs_golden(filter_mode, stack):
    g, batch, pos2 = _setup()
    model = schnet_from(g, DEV)
    loss, acc = do_InfoNCE(default_args("schnet", T=g.cfg["T"]), batch, model, None, 0.0, g.cfg["sigma"],
                           positions_02=pos2, stack_views=stack)
    assert rel_err(loss, g["out"]["infonce_loss"]) <= TOL_OUT and abs(acc - float(g["out"]["infonce_acc"])) < 1e-6
    loss.backward()
    _check(model, g, "infonce/", TOL_GRAD if filter_mode == "simt" else 1e-3)


@pytest.mark.parametrize("stack", [True, False])
def test_ebm_nce_vs_golden(filter_mode, stack):
    g, batch, pos2 = _setup()
    model = schnet_from(g, DEV)
    loss, acc = do_EBM_NCE(default_args("schnet"), batch, model, torch.nn.BCEWithLogitsLoss(), 0.0, g.cfg["sigma"],
                           num_neg=g.cfg["num_neg"], positions_02=pos2, stack_views=stack)
    assert loss.dtype == torch.float64 and rel_err(loss, g["out"]["ebm_loss"]) <= TOL_OUT
    assert abs(acc - float(g["out"]["ebm_acc"])) < 1e-6
    loss.backward()
    _check(model, g, "ebm/", TOL_GRAD if filter_mode == "simt" else 1e-3)


def test_distance_prediction_vs_golden(filter_mode):
    g, batch, _ = _setup()
    model = schnet_from(g, DEV)
    head = DistancePredictor(g.cfg["hidden"]).to(DEV)
    head.load_state_dict(g.sd("sdpred"), strict=True)
    loss = do_DistancePrediction(default_args("schnet"), batch, model, head)
    assert rel_err(loss, g["out"]["distance_loss"]) <= TOL_OUT
    loss.backward()
    tol = TOL_GRAD if filter_mode == "simt" else 1e-3
    _check(model, g, "distance/", tol)
    got = grads_of(head)
    for k, ref in g["grad"].items():
        if k.startswith("distance_predictor/"):
            kk = "predictor." + k[len("distance_predictor/"):]
            assert rel_err(got[kk], ref) <= tol, (k, rel_err(got[kk], ref))
    # the reference-shaped forward(u, v, d) gives the same loss as the fused pair path
    with torch.no_grad():
        _, h = model(batch.x[:, 0], batch.positions, batch.batch, return_latent=True)
        sei = batch.super_edge_index
        d = (batch.positions[sei[0]] - batch.positions[sei[1]]).norm(dim=1)
        assert rel_err(head(h[sei[0]], h[sei[1]], d), loss) <= 1e-5


def test_cycle_index():
    assert cycle_index(5, 2).tolist() == [2, 3, 4, 0, 1]
