"""GPU: the tcgen05 / TMEM plumbing (split-precision GEMM self test) and the tensor-core filter kernel."""
import ctypes

import pytest
import torch

from _golden import rel_err
from geossl_b200 import _lib, ops
from geossl_b200.data import synthetic_batch
from oracle import models as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _selftest(mode, fp16, a, b, K, N):
    d = torch.full((128, N), float("nan"), device=DEV)
    rc = _lib.load().geossl_tc_selftest(mode, fp16, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), K, N,
                                        ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "tc_selftest")
    torch.cuda.synchronize()
    return d


@pytest.mark.parametrize("fp16,tol", [(1, 2e-6), (0, 4e-5)])
@pytest.mark.parametrize("K", [64, 128])
def test_tc_selftest_k_major(fp16, tol, K):
    g = torch.Generator().manual_seed(K + fp16)
    a, b = torch.randn(128, K, generator=g), torch.randn(128, K, generator=g)
    d = _selftest(0, fp16, a.to(DEV), b.to(DEV), K, 128)
    ref = a.double() @ b.double().t()
    assert rel_err(d, ref) <= tol, rel_err(d, ref)


@pytest.mark.parametrize("fp16,tol", [(1, 2e-6), (0, 4e-5)])
@pytest.mark.parametrize("N", [64, 128])
def test_tc_selftest_mn_major(fp16, tol, N):
    g = torch.Generator().manual_seed(N + fp16)
    x, y = torch.randn(128, 128, generator=g), torch.randn(128, N, generator=g)
    d = _selftest(1, fp16, x.to(DEV), y.to(DEV), 128, N)
    ref = x.double().t() @ y.double()
    assert rel_err(d, ref) <= tol, rel_err(d, ref)


@pytest.mark.parametrize("fp16,tol", [(1, 2e-6), (0, 4e-5)])
@pytest.mark.parametrize("K,N", [(64, 128), (128, 128), (128, 64)])
def test_tc_selftest_a_in_tmem(fp16, tol, K, N):
    g = torch.Generator().manual_seed(K + N + fp16)
    a, b = torch.randn(128, K, generator=g), torch.randn(128, K, generator=g)     # only the first N rows of b are used
    d = _selftest(3, fp16, a.to(DEV), b.to(DEV), K, N)
    ref = a.double() @ b[:N].double().t()
    assert rel_err(d, ref) <= tol, rel_err(d, ref)


@pytest.mark.parametrize("mode,tol", [("tc_fp16", 3e-6), ("tc_bf16", 1e-4)])
@pytest.mark.parametrize("G,ng,lo,hi", [(50, 8, 20, 40), (64, 3, 5, 9), (20, 40, 25, 35)])
def test_filter_fwd_tc_vs_oracle(mode, tol, G, ng, lo, hi):
    b = synthetic_batch(ng, lo, hi, seed=G, with_pairs=False)
    cutoff = 10.0
    gen = torch.Generator().manual_seed(1)
    w1, b1 = torch.randn(128, G, generator=gen) * 0.3, torch.randn(128, generator=gen) * 0.1
    w2, b2 = torch.randn(128, 128, generator=gen) * 0.15, torch.randn(128, generator=gen) * 0.1
    offset = torch.linspace(0.0, cutoff, G)
    coeff = O.smearing_coeff(offset)
    ei = O.radius_graph(b.positions, cutoff, b.batch)
    d = (b.positions[ei[0]] - b.positions[ei[1]]).norm(dim=-1)
    sd = {"interactions.0.mlp.0.weight": w1, "interactions.0.mlp.0.bias": b1,
          "interactions.0.mlp.2.weight": w2, "interactions.0.mlp.2.bias": b2}
    W = O.schnet_filter(sd, 0, d, O.gaussian_smearing(d, offset), cutoff)
    graph = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), cutoff, num_graphs=ng)
    e = graph.num_edges
    dev = [t.to(DEV) for t in (w1, b1, w2, b2)]
    filt = ops.filter_forward(graph, offset.to(DEV), coeff, cutoff, *dev, mode=mode)
    torch.cuda.synchronize()
    assert rel_err(filt[:e], W) <= tol, rel_err(filt[:e], W)
    simt = ops.filter_forward(graph, offset.to(DEV), coeff, cutoff, *dev, mode="simt")
    assert rel_err(filt[:e], simt[:e]) <= tol


@pytest.mark.parametrize("share", [True, False])
@pytest.mark.parametrize("G,ng,lo,hi,density", [(50, 8, 20, 40, 0.05), (63, 3, 5, 9, 0.05), (20, 40, 25, 35, 0.05),
                                                (51, 300, 28, 32, 0.05), (50, 6, 40, 70, 0.1)])
def test_filter_bwd_tc_vs_simt(G, ng, lo, hi, density, share, monkeypatch):
    """Tensor-core backward (bf16 split, TMEM-resident weight gradients) against the exact fp32 per-edge kernel; with
    ``share`` the tensor-core path runs over undirected pairs (the last case has truncated rows => orphan edges)."""
    monkeypatch.setattr(ops, "SHARE_PAIR_FILTERS", share)
    b = synthetic_batch(ng, lo, hi, seed=G + 1, with_pairs=False, density=density)
    cutoff = 10.0
    gen = torch.Generator().manual_seed(3)
    params = [(torch.randn(128, G, generator=gen) * 0.3), torch.randn(128, generator=gen) * 0.1,
              torch.randn(128, 128, generator=gen) * 0.15, torch.randn(128, generator=gen) * 0.1]
    offset = torch.linspace(0.0, cutoff, G).to(DEV)
    coeff = O.smearing_coeff(offset.cpu())
    n = b.positions.shape[0]
    x, gout = torch.randn(n, 128, generator=gen), torch.randn(n, 128, generator=gen) * 1e-3
    graph = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), cutoff, num_graphs=ng)
    res = {}
    old = ops.FILTER_MODE
    try:
        for mode in ("simt", "tc_fp16"):
            ops.FILTER_MODE = mode
            leaves = [p.clone().to(DEV).requires_grad_() for p in params]
            xc = x.to(DEV).requires_grad_()
            out = ops.CFConvLayer.apply(xc, *leaves, offset, graph, coeff, cutoff)
            out.backward(gout.to(DEV))
            torch.cuda.synchronize()
            res[mode] = [xc.grad] + [p.grad for p in leaves]
    finally:
        ops.FILTER_MODE = old
    for name, a, ref in zip(("x", "w1", "b1", "w2", "b2"), res["tc_fp16"], res["simt"]):
        assert torch.isfinite(a).all(), name
        assert rel_err(a, ref) <= 5e-5, (name, rel_err(a, ref))


@pytest.mark.parametrize("n,pre_ssp,use_bias,use_res", [(7680, False, False, False), (1000, True, True, True), (63, False, True, False),
                                                         (129, True, True, False), (20000, True, True, True)])
def test_linear_tc_vs_torch(n, pre_ssp, use_bias, use_res):
    """128 -> 128 atom-wise layer on the tensor cores (fp16-split forward, bf16-split gradients) against fp64 torch."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(n)
    x = (torch.randn(n, 128, generator=g) * 2).to(DEV).requires_grad_()
    w = (torch.randn(128, 128, generator=g) * 0.1).to(DEV).requires_grad_()
    b = (torch.randn(128, generator=g) * 0.1).to(DEV).requires_grad_() if use_bias else None
    r = torch.randn(n, 128, generator=g).to(DEV).requires_grad_() if use_res else None
    gy = (torch.randn(n, 128, generator=g) * 1e-3).to(DEV)
    y = ops.LinearTC.apply(x, w, b, r, pre_ssp)
    y.backward(gy)
    xd, wd = x.detach().double().requires_grad_(), w.detach().double().requires_grad_()
    bd = b.detach().double().requires_grad_() if use_bias else None
    rd = r.detach().double().requires_grad_() if use_res else None
    inp = (F.softplus(xd) - 0.6931471805599453) if pre_ssp else xd
    yd = F.linear(inp, wd, bd)
    if use_res:
        yd = yd + rd
    yd.backward(gy.double())
    assert rel_err(y, yd) <= 3e-6, rel_err(y, yd)
    assert rel_err(x.grad, xd.grad) <= 5e-5 and rel_err(w.grad, wd.grad) <= 5e-5, (rel_err(x.grad, xd.grad), rel_err(w.grad, wd.grad))
    if use_bias:
        assert rel_err(b.grad, bd.grad) <= 1e-5
    if use_res:
        assert torch.equal(r.grad, gy)



@pytest.mark.parametrize("ng,lo,hi,density", [(8, 20, 40, 0.05), (6, 40, 70, 0.1), (5, 1, 3, 0.05)])
def test_pair_sharing_forward_is_bit_identical(ng, lo, hi, density, monkeypatch):
    """|pos_j - pos_i| == |pos_i - pos_j| bit for bit, so sharing one filter row per atom pair must not change a single
    bit of the aggregate m_i = sum_e x[src_e] * W_e (same kernel arithmetic, same summation order)."""
    b = synthetic_batch(ng, lo, hi, seed=7, with_pairs=False, density=density)
    gen = torch.Generator().manual_seed(5)
    G, cutoff = 50, 10.0
    params = [(torch.randn(128, G, generator=gen) * 0.3).to(DEV), (torch.randn(128, generator=gen) * 0.1).to(DEV),
              (torch.randn(128, 128, generator=gen) * 0.15).to(DEV), (torch.randn(128, generator=gen) * 0.1).to(DEV)]
    offset = torch.linspace(0.0, cutoff, G).to(DEV)
    coeff = O.smearing_coeff(offset.cpu())
    x = torch.randn(b.positions.shape[0], 128, generator=gen).to(DEV)
    out = {}
    for share in (False, True):
        monkeypatch.setattr(ops, "SHARE_PAIR_FILTERS", share)
        graph = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), cutoff, num_graphs=ng)
        out[share] = ops.CFConvLayer.apply(x, *params, offset, graph, coeff, cutoff)
        if share:
            e, u = graph.num_edges, int(graph.n_pairs_dev.item())
            assert e // 2 <= u <= e
    assert torch.equal(out[True], out[False])
