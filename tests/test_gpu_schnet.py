"""GPU: SchNet kernels and module against the oracle and the golden fixtures.
Tolerances are BASELINE.json's: representations rel 1e-5, parameter gradients rel 1e-4 (fp32 SIMT path),
rel = max|a-b| / max|b| on the tensor."""
import pytest
import torch
import torch.nn.functional as F

from _build import grads_of, schnet_from
from _golden import Golden, rel_err
from geossl_b200 import ops
from geossl_b200.data import synthetic_batch
from oracle import models as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_OUT, TOL_GRAD = 1e-5, 1e-4


def _layer_params(Fd, G, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(Fd, G, generator=g) * 0.3, torch.randn(Fd, generator=g) * 0.1,
            torch.randn(Fd, Fd, generator=g) * 0.15, torch.randn(Fd, generator=g) * 0.1)


@pytest.mark.parametrize("Fd,G,ng,lo,hi", [(128, 50, 6, 20, 40), (64, 51, 5, 5, 30), (32, 20, 4, 3, 12), (128, 64, 2, 70, 90)])
def test_filter_and_cfconv_kernels_vs_oracle(Fd, G, ng, lo, hi, filter_mode):
    b = synthetic_batch(ng, lo, hi, seed=Fd + G, with_pairs=False)
    cutoff = 10.0
    w1, b1, w2, b2 = _layer_params(Fd, G, 1)
    offset = torch.linspace(0.0, cutoff, G)
    coeff = O.smearing_coeff(offset)
    n = b.positions.shape[0]
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(n, Fd, generator=gen)
    gout = torch.randn(n, Fd, generator=gen)
    # oracle
    ei = O.radius_graph(b.positions, cutoff, b.batch)
    d = (b.positions[ei[0]] - b.positions[ei[1]]).norm(dim=-1)
    sd = {"interactions.0.mlp.0.weight": w1.clone().requires_grad_(), "interactions.0.mlp.0.bias": b1.clone().requires_grad_(),
          "interactions.0.mlp.2.weight": w2.clone().requires_grad_(), "interactions.0.mlp.2.bias": b2.clone().requires_grad_()}
    xr = x.clone().requires_grad_()
    W = O.schnet_filter(sd, 0, d, O.gaussian_smearing(d, offset), cutoff)
    m = O.cfconv_aggregate(xr, W, ei)
    (m * gout).sum().backward()
    # product
    graph = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), cutoff, num_graphs=ng)
    e = graph.num_edges
    assert e == ei.shape[1]
    cw = [t.to(DEV).requires_grad_() for t in (w1, b1, w2, b2)]
    filt = ops.filter_forward(graph, offset.to(DEV), coeff, cutoff, *[t.detach() for t in cw])
    assert rel_err(filt[:e], W) <= TOL_OUT
    xc = x.to(DEV).requires_grad_()
    out = ops.CFConvLayer.apply(xc, *cw, offset.to(DEV), graph, coeff, cutoff)
    assert rel_err(out, m) <= TOL_OUT
    out.backward(gout.to(DEV))
    assert rel_err(xc.grad, xr.grad) <= TOL_GRAD
    for got, key in zip(cw, sd):
        assert rel_err(got.grad, sd[key].grad) <= TOL_GRAD, key
    # the three composable primitives, including the materialised edge product
    ge = graph.exact()
    dW = ops.CFConvEdgeProduct.apply(x.to(DEV), gout.to(DEV), ge)
    assert rel_err(dW, x[ei[0]] * gout[ei[1]]) <= 1e-6
    assert rel_err(ops.CFConvAggregateT.apply(filt[:e].contiguous(), gout.to(DEV), ge), xr.grad) <= TOL_GRAD


def test_cfconv_is_deterministic_and_linear():
    b = synthetic_batch(64, 10, 60, seed=9, with_pairs=False)
    graph = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), 10.0, num_graphs=64)
    n, e = b.positions.shape[0], graph.num_edges
    x1, x2 = torch.randn(n, 128, device=DEV), torch.randn(n, 128, device=DEV)
    W = torch.randn(graph.capacity, 128, device=DEV)
    a = ops.CFConvAggregate.apply(x1, W, graph)
    assert torch.equal(a, ops.CFConvAggregate.apply(x1, W, graph))            # atomic free => bitwise repeatable
    lin = ops.CFConvAggregate.apply(x1 + 2 * x2, W, graph)
    assert rel_err(lin, a + 2 * ops.CFConvAggregate.apply(x2, W, graph)) <= 1e-5
    # adjoint identity <A x, g> = <x, A^T g>
    gq = torch.randn(n, 128, device=DEV)
    lhs = (a.double() * gq.double()).sum()
    rhs = (x1.double() * ops.CFConvAggregateT.apply(W, gq, graph).double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


@pytest.mark.parametrize("ng,lo,hi,density,n_max,owner_small", [
    (64, 10, 60, 0.05, 60, False),        # Molecule3D-like sizes
    (37, 25, 31, 0.05, 31, True),
    (6, 50, 64, 0.3, 64, False),          # dense: rows truncated at 32 neighbours => orphan pairs (one direction only)
    (5, 40, 70, 0.3, 48, False),          # graphs above the bound take the in-kernel global-memory path
    (3, 1, 2, 0.05, 8, False),            # single atoms / one pair
])
def test_pair_centric_cfconv_matches_row_gather(ng, lo, hi, density, n_max, owner_small, monkeypatch):
    """geossl_cfconv_pairs (one CTA per graph, every filter row read once, both endpoints updated) against the row-gather
    kernels with shared filter rows: same sums in another fp32 order.  Covers orphan pairs, the oversize-graph path,
    both adjoint directions, every tuning code, and bitwise repeatability (no atomics)."""
    monkeypatch.setattr(ops, "PAIR_OWNER_SMALL", owner_small)
    b = synthetic_batch(ng, lo, hi, seed=3 * ng + lo, with_pairs=False, density=density)
    graph = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), 10.0, num_graphs=ng).ensure_pairs()
    n = b.positions.shape[0]
    u, e = int(graph.n_pairs_dev.item()), graph.num_edges
    if density > 0.2:
        assert u > e // 2                                            # some pair lost its reverse direction
    gen = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(n, 128, device=DEV, generator=gen)
    W = torch.randn(graph.capacity, 128, device=DEV, generator=gen)
    want_f = ops._cfconv_fwd(x, W, graph, graph.pair_of_edge)
    want_t = ops._cfconv_bwd_x(W, x, graph, graph.pair_of_edge)
    graph.max_graph_atoms = n_max
    for tuning in (0, 116, 208, 216, 308, 316, 408, 416):
        monkeypatch.setattr(ops, "CFCONV_PAIRS_TUNING", tuning)
        got_f = ops._cfconv_pairs(x, W, graph, False)
        got_t = ops._cfconv_pairs(x, W, graph, True)
        assert rel_err(got_f, want_f) <= 2e-6, tuning
        assert rel_err(got_t, want_t) <= 2e-6, tuning
        assert torch.equal(got_f, ops._cfconv_pairs(x, W, graph, False))
    # adjoint identity <A x, g> = <x, A^T g> between the two directions of the new kernel
    monkeypatch.setattr(ops, "CFCONV_PAIRS_TUNING", 0)
    gq = torch.randn(n, 128, device=DEV, generator=gen)
    lhs = (ops._cfconv_pairs(x, W, graph, False).double() * gq.double()).sum()
    rhs = (x.double() * ops._cfconv_pairs(gq, W, graph, True).double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


def test_fused_layer_takes_the_pair_centric_kernel_when_graph_sizes_are_known():
    """CFConvLayer forward + backward with and without the host-side bound on the graph size: identical results up to
    summation order; the bound is what selects geossl_cfconv_pairs (no source-sorted view is built then)."""
    b = synthetic_batch(16, 12, 30, seed=4, with_pairs=False)
    assert b.extras["max_graph_atoms"] == int(torch.bincount(b.batch).max())
    w1, b1, w2, b2 = _layer_params(128, 50, 1)
    offset = torch.linspace(0.0, 10.0, 50)
    coeff = O.smearing_coeff(offset)
    gen = torch.Generator().manual_seed(2)
    x, gout = torch.randn(b.positions.shape[0], 128, generator=gen), torch.randn(b.positions.shape[0], 128, generator=gen)
    res = []
    for bound in (None, b.extras["max_graph_atoms"]):
        graph = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), 10.0, num_graphs=16, max_graph_atoms=bound)
        if ops.FILTER_MODE != "simt":
            assert (graph.t_rowptr is None) == (bound is not None)
        cw = [t.to(DEV).requires_grad_() for t in (w1, b1, w2, b2)]
        xc = x.to(DEV).requires_grad_()
        out = ops.CFConvLayer.apply(xc, *cw, offset.to(DEV), graph, coeff, 10.0)
        out.backward(gout.to(DEV))
        res.append([out.detach(), xc.grad] + [t.grad for t in cw])
    for a, c in zip(*res):
        assert rel_err(a, c) <= 2e-6


@pytest.mark.parametrize("name", ["schnet_small", "schnet_trunc"])
def test_schnet_module_vs_golden(name, filter_mode):
    g = Golden(name)
    m = schnet_from(g, DEV)
    i = g["in"]
    out, h = m(i["z"].to(DEV), i["pos"].to(DEV), i["batch"].to(DEV), return_latent=True)
    assert rel_err(h, g["out"]["h"]) <= TOL_OUT and rel_err(out, g["out"]["out"]) <= TOL_OUT
    ((h * i["w_h"].to(DEV)).sum() + (out * i["w_o"].to(DEV)).sum()).backward()
    got = grads_of(m)
    for k, ref in g["grad"].items():
        assert rel_err(got[k], ref) <= TOL_GRAD, (k, rel_err(got[k], ref))


def test_schnet_standalone_block_api():
    """InteractionBlock / CFConv keep the reference signature (x, edge_index, edge_weight, edge_attr)."""
    g = Golden("schnet_small")
    m = schnet_from(g, DEV)
    i = g["in"]
    pos, batch = i["pos"].to(DEV), i["batch"].to(DEV)
    ei = ops.radius_graph(pos, g.cfg["cutoff"], batch)
    ew = (pos[ei[0]] - pos[ei[1]]).norm(dim=-1)
    ea = m.distance_expansion(ew)
    h = m.embedding(i["z"].to(DEV))
    y = m.interactions[0](h, ei, ew, ea, batch)
    sd = g.sd()
    W = O.schnet_filter(sd, 0, ew.cpu(), ea.cpu(), g.cfg["cutoff"])
    x = F.linear(h.cpu(), sd["interactions.0.conv.lin1.weight"])
    x = O.cfconv_aggregate(x, W, ei.cpu())
    x = F.linear(x, sd["interactions.0.conv.lin2.weight"], sd["interactions.0.conv.lin2.bias"])
    x = F.linear(O.shifted_softplus(x), sd["interactions.0.lin.weight"], sd["interactions.0.lin.bias"])
    assert rel_err(y, x.detach()) <= TOL_OUT


def test_md17_double_backward_vs_golden():
    """Energy + autograd force + backward through the force (finetune_md17.py:32-54)."""
    g = Golden("md17_small")
    m = schnet_from(g, DEV)
    lin = torch.nn.Linear(g.cfg["hidden"], 1).to(DEV)
    lin.load_state_dict(g.sd("sdlin", DEV))
    i = g["in"]
    pos = i["pos"].to(DEV).requires_grad_()
    rep = m(i["z"].to(DEV), pos, i["batch"].to(DEV))
    energy = lin(rep).squeeze(1)
    force = -torch.autograd.grad(energy, pos, grad_outputs=torch.ones_like(energy), create_graph=True, retain_graph=True)[0]
    loss = 0.05 * F.l1_loss(energy, i["y"].to(DEV)) + 0.95 * F.l1_loss(force, i["force_target"].to(DEV))
    loss.backward()
    assert rel_err(energy, g["out"]["energy"]) <= TOL_OUT and rel_err(force, g["out"]["force"]) <= TOL_GRAD
    assert rel_err(loss, g["out"]["loss"]) <= TOL_OUT
    got = grads_of(m)
    for k, ref in g["grad"].items():
        assert rel_err(got[k], ref) <= 2e-4, (k, rel_err(got[k], ref))
    for k, ref in g["gradlin"].items():
        assert rel_err(dict(lin.named_parameters())[k].grad, ref) <= 2e-4


def test_md17_force_step_at_full_width_vs_oracle(filter_mode):
    """finetune_md17.py:32-54 at the real model width (hidden = filters = 128, 50 gaussians): energy, autograd force and the
    parameter gradients of the force loss (double backward) against the CPU oracle computed in the test.  In the
    tensor-core modes the edge-sized filter MLP runs on ops.MatXWt / MatXW / MatTX (tcgen05 blocks closed under
    differentiation) instead of library GEMMs; the 32-wide golden fixture ``md17_small`` cannot reach that path."""
    from geossl_b200.Geom3D.models import SchNet
    torch.manual_seed(3)
    m = SchNet(hidden_channels=128, num_filters=128, num_interactions=3, num_gaussians=50, cutoff=10.0, node_class=9, readout="mean")
    lin = torch.nn.Linear(128, 1)
    b = synthetic_batch(6, 21, seed=8, with_pairs=False)
    gen = torch.Generator().manual_seed(4)
    y, ft = torch.randn(6, generator=gen), torch.randn(b.positions.shape, generator=gen)
    # oracle (CPU, torch fp32)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and v.dtype == torch.float32) for k, v in m.state_dict().items()}
    sdl = {k: v.clone().requires_grad_() for k, v in lin.state_dict().items()}
    pos = b.positions.clone().requires_grad_()
    out, _ = O.schnet_forward(sd, b.x[:, 0].contiguous(), pos, b.batch, cutoff=10.0, readout="mean")
    e_ref = F.linear(out, sdl["weight"], sdl["bias"]).squeeze(1)
    f_ref = -torch.autograd.grad(e_ref, pos, grad_outputs=torch.ones_like(e_ref), create_graph=True, retain_graph=True)[0]
    l_ref = 0.05 * F.l1_loss(e_ref, y) + 0.95 * F.l1_loss(f_ref, ft)
    l_ref.backward()
    # product
    m.to(DEV), lin.to(DEV)
    used = []
    orig = ops.filter_mlp
    ops.filter_mlp = lambda *a: used.append(1) or orig(*a)
    try:
        posd = b.positions.to(DEV).requires_grad_()
        rep = m(b.x[:, 0].contiguous().to(DEV), posd, b.batch.to(DEV), num_graphs=6)
        energy = lin(rep).squeeze(1)
        force = -torch.autograd.grad(energy, posd, grad_outputs=torch.ones_like(energy), create_graph=True, retain_graph=True)[0]
        loss = 0.05 * F.l1_loss(energy, y.to(DEV)) + 0.95 * F.l1_loss(force, ft.to(DEV))
        loss.backward()
    finally:
        ops.filter_mlp = orig
    assert bool(used) == (filter_mode != "simt")
    assert rel_err(energy, e_ref) <= TOL_OUT and rel_err(force, f_ref) <= TOL_GRAD, (rel_err(energy, e_ref), rel_err(force, f_ref))
    assert rel_err(loss, l_ref) <= TOL_OUT
    for k, g in grads_of(m).items():
        if ".conv.nn." not in k:
            assert rel_err(g, sd[k].grad) <= 2e-4, (k, rel_err(g, sd[k].grad))
    for k, g in grads_of(lin).items():
        assert rel_err(g, sdl[k].grad) <= 2e-4, (k, rel_err(g, sdl[k].grad))


@pytest.mark.parametrize("density", [0.05, 0.3])
def test_pair_row_primitives_match_per_edge_primitives_to_second_order(density):
    """CFConvAggregateP / TP / PairProduct (one filter row per atom pair) against the per-edge trio fed with
    W_e = W_u[pair_of_edge]: value, first derivatives and a second derivative.  density 0.3 truncates rows at 32
    neighbours, so some pairs exist in one direction only."""
    b = synthetic_batch(5, 30, 60, seed=12, with_pairs=False, density=density)
    ge = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), 10.0, num_graphs=5).exact().ensure_pairs()
    n, u = b.positions.shape[0], ge.num_pairs
    gen = torch.Generator(device=DEV).manual_seed(1)
    x0 = torch.randn(n, 128, device=DEV, generator=gen)
    w0 = torch.randn(u, 128, device=DEV, generator=gen)
    c = torch.randn(n, 128, device=DEV, generator=gen)
    poe = ge.pair_of_edge[:ge.num_edges].long()

    def run(agg):
        x, w = x0.clone().requires_grad_(), w0.clone().requires_grad_()
        m = agg(x, w)
        gx, gw = torch.autograd.grad((torch.tanh(m) * c).sum(), (x, w), create_graph=True)
        second = torch.autograd.grad(gx.square().sum() + gw.square().sum(), (x, w))
        return m, gx, gw, second[0], second[1]

    got = run(lambda x, w: ops.CFConvAggregateP.apply(x, w, ge))
    want = run(lambda x, w: ops.CFConvAggregate.apply(x, w[poe], ge))
    for a, r in zip(got, want):
        assert rel_err(a, r) <= 1e-5, rel_err(a, r)


def test_closed_elementwise_families_match_torch_to_third_order():
    """ops.SspFamily (shifted softplus and its derivatives) and ops.RowScale / RowDot against torch's own formulas in float64:
    value, gradient, gradient of a gradient norm, and one more order for the activation (incl. inputs above torch's
    softplus threshold of 20, where the derivative is exactly 1)."""
    gen = torch.Generator(device=DEV).manual_seed(9)
    a = 6 * torch.randn(777, 128, device=DEV, generator=gen)
    a[0, :4] = torch.tensor([25.0, 20.5, -30.0, 0.0], device=DEV)
    c = torch.randn(777, device=DEV, generator=gen)
    w = torch.randn(777, 128, device=DEV, generator=gen)

    def chain(av, cv, act, scale):
        y = scale(act(av), cv)
        g1a, g1c = torch.autograd.grad((y * w.to(y.dtype)).sum(), (av, cv), create_graph=True)
        g2a, g2c = torch.autograd.grad(g1a.square().sum() + g1c.square().sum(), (av, cv), create_graph=True)
        g3a, = torch.autograd.grad(g2a.square().sum(), av)
        return y, g1a, g1c, g2a, g2c, g3a

    got = chain(a.clone().requires_grad_(), c.clone().requires_grad_(), ops.ssp_any_order, ops.row_scale)
    want = chain(a.double().requires_grad_(), c.double().requires_grad_(),
                 lambda t: F.softplus(t) - 0.6931471805599453, lambda t, cv: t * cv.view(-1, 1))
    for g, r in zip(got, want):
        assert rel_err(g, r.float()) <= 2e-5, rel_err(g, r.float())


def test_tensor_core_products_are_closed_under_differentiation():
    """ops.MatXWt / MatXW / MatTX against torch matmul in float64: values, first derivatives and a second derivative
    (gradient of a gradient norm), with a K-padded operand (64 live columns)."""
    gen = torch.Generator(device=DEV).manual_seed(6)
    n = 1000
    x = torch.randn(n, 64, device=DEV, generator=gen).requires_grad_()
    w = (0.2 * torch.randn(128, 128, device=DEV, generator=gen))
    w[:, 64:] = 0
    w.requires_grad_()
    bias = torch.randn(128, device=DEV, generator=gen).requires_grad_()
    x64, w64, b64 = (t.detach().double().requires_grad_() for t in (x, w, bias))

    def second(xv, wv, bv, mm):
        yv = torch.tanh(mm(xv, wv, bv))
        gx, = torch.autograd.grad(yv.square().sum(), xv, create_graph=True)
        return yv, gx, torch.autograd.grad(gx.square().sum(), (wv, bv))

    y, gx, (gw2, gb2) = second(x, w, bias, lambda a, b_, c: ops.MatXWt.apply(a, b_, c, True))
    yr, gxr, (gw2r, gb2r) = second(x64, w64, b64, lambda a, b_, c: a @ b_[:, :64].t() + c)
    assert rel_err(y, yr.float()) <= 1e-5 and rel_err(gx, gxr.float()) <= 1e-4
    assert rel_err(gw2[:, :64], gw2r[:, :64].float()) <= 2e-4 and rel_err(gb2, gb2r.float()) <= 2e-4


def test_full_size_forward_properties_config2(filter_mode):
    """Config 2 (256 x 30 atoms, H=F=128, G=50, L=6): permuting whole molecules permutes the result."""
    torch.manual_seed(0)
    from geossl_b200.Geom3D.models import SchNet
    m = SchNet(node_class=9).to(DEV)
    b = synthetic_batch(256, 30, seed=4, with_pairs=False).to(DEV)
    z = b.x[:, 0].contiguous()
    with torch.no_grad():
        out, h = m(z, b.positions, b.batch, return_latent=True, num_graphs=256)
        perm = torch.arange(256, device=DEV).flip(0)
        idx = (perm[:, None] * 30 + torch.arange(30, device=DEV)[None]).reshape(-1)
        out2, h2 = m(z[idx], b.positions[idx], b.batch, return_latent=True, num_graphs=256)
    assert torch.isfinite(h).all()
    assert rel_err(h2, h[idx]) <= 1e-6 and rel_err(out2, out[perm]) <= 1e-6


def test_lba_shaped_pockets_vs_oracle(filter_mode):
    """BASELINE configs[4] shape: ~600-atom pockets, cutoff 6 A (every row truncates to 32/33 neighbours)."""
    torch.manual_seed(3)
    from geossl_b200.Geom3D.models import SchNet
    m = SchNet(hidden_channels=128, num_filters=128, num_interactions=2, num_gaussians=50, cutoff=6.0, node_class=9, readout="mean")
    b = synthetic_batch(2, 520, 600, seed=8, with_pairs=False)
    z = b.x[:, 0].contiguous()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    out_ref, h_ref, ei = O.schnet_forward(sd, z, b.positions, b.batch, cutoff=6.0, readout="mean", return_edge_index=True)
    m.to(DEV)
    out, h = m(z.to(DEV), b.positions.to(DEV), b.batch.to(DEV), return_latent=True)
    deg = torch.bincount(ei[1], minlength=z.numel())
    assert int(deg.max()) == 33 and float((deg >= 32).float().mean()) > 0.5
    assert rel_err(h, h_ref) <= TOL_OUT and rel_err(out, out_ref) <= TOL_OUT


def test_md17_train_step_through_finetune_module():
    """The fine-tune caller (finetune_md17.py::train) via geossl_b200.finetune: same losses as the golden, and one Adam
    step changes the weights."""
    from geossl_b200.data import AtomTupleBatch
    from geossl_b200.finetune import md17_losses, md17_train_step
    from geossl_b200.pretrain import default_args
    g = Golden("md17_small")
    m = schnet_from(g, DEV)
    lin = torch.nn.Linear(g.cfg["hidden"], 1).to(DEV)
    lin.load_state_dict(g.sd("sdlin", DEV))
    i = g["in"]
    z = i["z"].to(DEV)
    batch = AtomTupleBatch(torch.stack([z, torch.zeros_like(z)], 1), i["pos"].to(DEV), i["batch"].to(DEV), None,
                           n_graphs=int(i["batch"][-1]) + 1, extras=dict(y=i["y"].to(DEV), force=i["force_target"].to(DEV)))
    crit = torch.nn.L1Loss()
    loss, energy, force = md17_losses(default_args("schnet"), batch, m, lin, crit)
    assert rel_err(loss, g["out"]["loss"]) <= TOL_OUT and rel_err(force, g["out"]["force"]) <= TOL_GRAD
    opt = torch.optim.Adam(list(m.parameters()) + list(lin.parameters()), lr=1e-3)
    before = m.lin1.weight.detach().clone()
    batch.positions = i["pos"].to(DEV)
    l2 = md17_train_step(default_args("schnet"), batch, m, lin, crit, opt)
    assert rel_err(l2, g["out"]["loss"]) <= TOL_OUT and not torch.equal(before, m.lin1.weight)


def _lba_case(n_graphs, layers, seed):
    from geossl_b200.Geom3D.models import SchNet
    torch.manual_seed(seed)
    m = SchNet(hidden_channels=128, num_filters=128, num_interactions=layers, num_gaussians=50, cutoff=6.0, node_class=9, readout="mean")
    lin = torch.nn.Linear(128, 1)
    b = synthetic_batch(n_graphs, 560, 640, seed=seed + 1, with_pairs=False)
    b.extras["y"] = torch.randn(n_graphs, generator=torch.Generator().manual_seed(seed + 2))
    return m, lin, b


def _oracle_lba(m, lin, b):
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and v.dtype == torch.float32) for k, v in m.state_dict().items()}
    sdl = {k: v.clone().requires_grad_() for k, v in lin.state_dict().items()}
    out, _ = O.schnet_forward(sd, b.x[:, 0].contiguous(), b.positions, b.batch, cutoff=6.0, readout="mean")
    loss = F.mse_loss(F.linear(out, sdl["weight"], sdl["bias"]).squeeze(), b.extras["y"])      # finetune_lba.py:42-47,244
    loss.backward()
    return loss, sd, sdl


def test_lba_loss_and_backward_at_pocket_shape(filter_mode):
    """BASELINE configs[4] shape (600-atom pockets, cutoff 6 A, full 6-layer model): ``finetune.lba_loss`` (the body of
    finetune_lba.py::train :33-47) and EVERY parameter gradient against the CPU oracle.  Rows truncate at 32/33
    neighbours, so the graph is asymmetric and most filter rows are 'orphan' pairs."""
    from geossl_b200.finetune import lba_loss
    from geossl_b200.pretrain import default_args
    m, lin, b = _lba_case(4, 6, 11)
    ref, sd, sdl = _oracle_lba(m, lin, b)
    m.to(DEV), lin.to(DEV)
    loss = lba_loss(default_args("schnet"), b.to(DEV), m, lin, torch.nn.MSELoss())
    assert rel_err(loss, ref) <= TOL_OUT, rel_err(loss, ref)
    loss.backward()
    tol = TOL_GRAD if filter_mode == "simt" else 2e-4            # stated tensor-core bound (tests/test_gpu_ddm.py)
    for k, g in grads_of(m).items():
        if ".conv.nn." not in k:
            assert rel_err(g, sd[k].grad) <= tol, (k, rel_err(g, sd[k].grad))
    for k, g in grads_of(lin).items():
        assert rel_err(g, sdl[k].grad) <= tol, (k, rel_err(g, sdl[k].grad))


def test_graphed_finetune_step_pads_pockets_of_different_size():
    """finetune.GraphedFinetuneStep: one captured graph serves pockets of different atom counts (padding atoms form an
    extra graph whose readout row is dropped); the replayed loss equals the eager loss of the same batch."""
    from geossl_b200.finetune import GraphedFinetuneStep, lba_loss, lba_train_step
    from geossl_b200.pretrain import default_args
    m, lin, _ = _lba_case(3, 2, 21)
    m.to(DEV), lin.to(DEV)
    pool = []
    for s in range(4):
        b = synthetic_batch(3, 100, 160, seed=40 + s, with_pairs=False)
        b.extras["y"] = torch.randn(3, generator=torch.Generator().manual_seed(s))
        pool.append(b.to(DEV))
    assert len({b.positions.size(0) for b in pool}) > 1
    crit = torch.nn.MSELoss()
    opt = torch.optim.Adam(list(m.parameters()) + list(lin.parameters()), lr=0.0, fused=True, capturable=True)   # lr 0: weights stay put
    targs = default_args("schnet")
    step = GraphedFinetuneStep(lambda b: lba_train_step(targs, b, m, lin, crit, opt, zero_grad=False), pool, opt)
    for b in pool:
        with torch.no_grad():
            ref = lba_loss(targs, b, m, lin, crit)
        assert rel_err(step(b), ref) <= 1e-6


def test_torch_custom_ops_match_the_autograd_layer():
    """torch.ops.geossl_b200.{radius_csr, cfconv, cfconv_transpose, linear128, pair_distance} call the same C-ABI entry points
    as geossl_b200.ops: identical results."""
    from geossl_b200 import torch_ops  # noqa: F401
    b = synthetic_batch(5, 20, 40, seed=4)
    pos, bt = b.positions.to(DEV), b.batch.to(DEV)
    g = ops.radius_csr(pos, bt, 10.0, num_graphs=5)
    rowptr, src, tgt, dist = torch.ops.geossl_b200.radius_csr(pos, bt, 10.0, 32, 5)
    e = g.num_edges
    assert torch.equal(rowptr, g.rowptr) and torch.equal(src[:e], g.src[:e]) and torch.equal(dist[:e], g.dist[:e])
    assert torch.equal(torch.ops.geossl_b200.radius_graph(pos, bt, 10.0, 32), g.edge_index)
    gen = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(pos.size(0), 128, device=DEV, generator=gen)
    filt = torch.randn(g.capacity, 128, device=DEV, generator=gen)
    assert torch.equal(torch.ops.geossl_b200.cfconv(x, filt, g.rowptr, g.src), ops.CFConvAggregate.apply(x, filt, g))
    assert torch.equal(torch.ops.geossl_b200.cfconv_transpose(filt, x, g.t_rowptr, g.t_eid, g.t_tgt),
                       ops.CFConvAggregateT.apply(filt, x, g))
    lin = torch.nn.Linear(128, 128).to(DEV)
    y = torch.ops.geossl_b200.linear128(x, lin.weight, lin.bias, True, x)
    assert rel_err(y, x + F.linear(F.softplus(x) - 0.6931471824645996, lin.weight, lin.bias)) <= 1e-5
    assert torch.equal(torch.ops.geossl_b200.pair_distance(pos, b.super_edge_index.to(DEV)), ops.pair_distance(pos, b.super_edge_index.to(DEV)))


def test_graphed_md17_step_equals_eager_loss():
    """finetune.GraphedMD17Step at the bench shape (256 aspirin-size conformers, full model): neighbour search + edge count
    eager, energy / autograd force / double backward / optimizer replayed from a CUDA graph.  With lr = 0 the replayed
    loss of every batch equals the eager loss of the same batch."""
    from geossl_b200.Geom3D.models import SchNet
    from geossl_b200.finetune import GraphedMD17Step, md17_losses, md17_train_step
    from geossl_b200.pretrain import default_args
    torch.manual_seed(5)
    m = SchNet(hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=10.0, node_class=9).to(DEV)
    lin = torch.nn.Linear(128, 1).to(DEV)
    crit = torch.nn.L1Loss()
    opt = torch.optim.Adam(list(m.parameters()) + list(lin.parameters()), lr=0.0, fused=True, capturable=True)

    def make(seed):
        b = synthetic_batch(256, 21, seed=seed, density=0.12, with_pairs=False)       # complete graphs (5.6 A cube)
        g = torch.Generator().manual_seed(seed)
        b.extras["y"] = torch.randn(256, generator=g)
        b.extras["force"] = torch.randn(b.positions.shape, generator=g)
        return b.to(DEV)

    pool = [make(s) for s in range(3)]
    targs = default_args("schnet")
    for b in pool:                                                   # as a training script would: a few eager iterations first
        md17_train_step(targs, b, m, lin, crit, opt)
    torch.cuda.synchronize()
    step = GraphedMD17Step(targs, pool[0], m, lin, crit, opt)
    for b in pool:
        got = step(b).clone()
        ref, _, _ = md17_losses(targs, b, m, lin, crit)
        assert rel_err(got, ref) <= 1e-6, rel_err(got, ref)
        del ref
    assert step.capture_error is None, step.capture_error
    assert len(step.graphs) == 1


@pytest.mark.parametrize("n_graphs,lo,hi", [(7, 5, 40), (1, 1, 1), (40, 30, 30)])
def test_fused_dense_chain_equals_layer_by_layer(n_graphs, lo, hi):
    """geossl_linear_chain_tc (conv.lin2 -> ssp -> lin + residual -> next conv.lin1 and the head in one launch each, forward
    and data-gradient chains) against one launch per layer: same representations and parameter gradients to fp32 round-off of
    the split-precision MMAs, including ragged last tiles and a single-atom graph."""
    from geossl_b200.Geom3D.models import SchNet
    torch.manual_seed(3)
    m = SchNet(hidden_channels=128, num_filters=128, num_interactions=3, num_gaussians=50, cutoff=10.0, node_class=9).to(DEV)
    b = synthetic_batch(n_graphs, lo, hi if hi > lo else None, seed=2, with_pairs=False).to(DEV)
    w = torch.randn(b.positions.size(0), 128, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    res = {}
    old = ops.FUSE_DENSE_CHAIN
    try:
        for fused in (True, False):
            ops.FUSE_DENSE_CHAIN = fused
            m.zero_grad(set_to_none=True)
            out, h = m(b.x[:, 0].contiguous(), b.positions, b.batch, return_latent=True, num_graphs=n_graphs)
            ((h * w).sum() + out.sum()).backward()
            res[fused] = (h.detach().clone(), {k: v.clone() for k, v in grads_of(m).items()})
    finally:
        ops.FUSE_DENSE_CHAIN = old
    assert rel_err(res[True][0], res[False][0]) <= 2e-6
    for k in res[False][1]:
        assert rel_err(res[True][1][k], res[False][1][k]) <= 2e-5, (k, rel_err(res[True][1][k], res[False][1][k]))
