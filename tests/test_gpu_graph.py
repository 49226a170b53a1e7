"""GPU: neighbour search / CSR kernels against oracle/radius.py -- bit-exact (integer work)."""
import numpy as np
import pytest
import torch

from _golden import Golden
from geossl_b200 import ops
from geossl_b200.data import synthetic_batch
from oracle.radius import radius_graph as oracle_radius_graph, radius_neighbors, transpose_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def check_graph(pos, batch, r, num_graphs=None, max_nb=32):
    g = ops.radius_csr(pos.to(DEV), batch.to(DEV), r, max_nb, num_graphs=num_graphs)
    rowptr, src = radius_neighbors(pos, r, batch, max_nb)
    e = int(rowptr[-1])
    assert g.num_edges == e
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr.astype(np.int32))
    assert np.array_equal(g.src[:e].cpu().numpy(), src.astype(np.int32))
    ei = g.edge_index.cpu()
    assert torch.equal(ei, oracle_radius_graph(pos, r, batch, max_num_neighbors=max_nb))
    t_rowptr, t_eid, t_tgt = transpose_csr(rowptr, src)
    assert np.array_equal(g.t_rowptr.cpu().numpy(), t_rowptr.astype(np.int32))
    assert np.array_equal(g.t_eid[:e].cpu().numpy(), t_eid.astype(np.int32))
    assert np.array_equal(g.t_tgt[:e].cpu().numpy(), t_tgt.astype(np.int32))
    if e:
        d = (pos[ei[0]] - pos[ei[1]]).norm(dim=-1)
        assert torch.allclose(g.dist[:e].cpu(), d, rtol=1e-6, atol=1e-7)
    return g


@pytest.mark.parametrize("name", ["schnet_small", "schnet_trunc"])
def test_radius_graph_matches_golden_edge_index(name):
    g = Golden(name)
    ei = ops.radius_graph(g["in"]["pos"].to(DEV), g.cfg["cutoff"], g["in"]["batch"].to(DEV))
    assert ei.dtype == torch.int64 and torch.equal(ei.cpu(), g["out"]["edge_index"])


@pytest.mark.parametrize("seed,ng,lo,hi,r,density", [
    (0, 7, 1, 9, 10.0, 0.05), (1, 16, 10, 60, 10.0, 0.05), (2, 5, 40, 90, 10.0, 0.08), (3, 6, 25, 70, 3.0, 0.05),
    (4, 3, 200, 600, 6.0, 0.05), (5, 64, 30, 30, 10.0, 0.05), (6, 4, 33, 35, 50.0, 0.05)])
def test_radius_csr_random(seed, ng, lo, hi, r, density):
    b = synthetic_batch(ng, lo, hi, seed=seed, density=density, with_pairs=False)
    g = check_graph(b.positions, b.batch, r, num_graphs=ng)
    deg = (g.rowptr[1:] - g.rowptr[:-1]).cpu()
    assert int(deg.max()) <= 33


def test_truncation_rows_of_32_and_33():
    b = synthetic_batch(3, 50, 64, seed=11, density=0.1, with_pairs=False)
    g = check_graph(b.positions, b.batch, 10.0)
    deg = (g.rowptr[1:] - g.rowptr[:-1]).cpu()
    assert int(deg.max()) == 33 and bool((deg == 32).any())     # asymmetric truncation


def test_boundary_ties_duplicates_and_tiny_graphs():
    # atoms exactly at distance r (strict <), duplicated positions (d = 0 kept, self dropped), 1-atom and empty graphs
    pos = torch.tensor([[0, 0, 0], [3, 4, 0], [0, 0, 0], [0, 0, 5], [0, 5.0000005, 0],
                        [1, 1, 1],
                        [2, 2, 2], [2, 2, 2]], dtype=torch.float32)
    batch = torch.tensor([0, 0, 0, 0, 0, 1, 3, 3])          # graph 2 is empty
    g = check_graph(pos, batch, 5.0, num_graphs=4)
    ei = g.edge_index.cpu()
    pairs = set(map(tuple, ei.t().tolist()))
    assert (2, 0) in pairs and (0, 2) in pairs and (1, 0) not in pairs and (3, 0) not in pairs
    assert (7, 6) in pairs and (6, 7) in pairs and not any(5 in p for p in pairs)


def test_small_max_num_neighbors_and_single_graph_default_batch():
    b = synthetic_batch(1, 40, seed=5, with_pairs=False)
    check_graph(b.positions, b.batch, 10.0, max_nb=4)
    ei = ops.radius_graph(b.positions.to(DEV), 4.0)             # batch=None
    assert torch.equal(ei.cpu(), oracle_radius_graph(b.positions, 4.0))


def test_empty_input():
    g = ops.radius_csr(torch.zeros((0, 3), device=DEV), torch.zeros(0, dtype=torch.long, device=DEV), 5.0, num_graphs=0)
    assert g.num_edges == 0 and g.edge_index.shape == (2, 0)


def test_full_size_properties_config2():
    """Config-2 sized batch (256 x U{10..60}) : size-independent invariants instead of the O(n^2) oracle."""
    b = synthetic_batch(256, 10, 60, seed=21, with_pairs=False)
    g = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), 10.0, num_graphs=256)
    e = g.num_edges
    rowptr, src, tgt = g.rowptr.cpu().long(), g.src[:e].cpu().long(), g.tgt[:e].cpu().long()
    deg = rowptr[1:] - rowptr[:-1]
    assert int(deg.max()) <= 33 and int(rowptr[-1]) == e
    assert bool((tgt[1:] >= tgt[:-1]).all())                                        # target-major
    same = tgt[1:] == tgt[:-1]
    assert bool((src[1:][same] > src[:-1][same]).all())                             # sources ascending, no duplicates
    assert bool((b.batch[src] == b.batch[tgt]).all()) and bool((src != tgt).all())  # intra-graph, no self loops
    d2 = ((b.positions[src] - b.positions[tgt]) ** 2).sum(-1)
    assert bool((d2 < 100.0 + 1e-3).all())
    untrunc = deg < 32                                                              # untruncated rows are complete
    n_in = torch.zeros(len(deg), dtype=torch.long)
    for gi in range(256):
        lo, hi = int(b.graph_ptr[gi]), int(b.graph_ptr[gi + 1])
        p = b.positions[lo:hi]
        dd = ((p[:, None] - p[None]) ** 2).sum(-1)
        n_in[lo:hi] = (dd < 100.0).sum(1) - 1
    assert bool((deg[untrunc] == n_in[untrunc]).all())
    # the transpose is a permutation of the edge ids grouped by source
    t_eid = g.t_eid[:e].cpu().long()
    assert torch.equal(torch.sort(t_eid).values, torch.arange(e))
    assert bool((src[t_eid][1:] >= src[t_eid][:-1]).all()) and torch.equal(g.t_tgt[:e].cpu().long(), tgt[t_eid])


@pytest.mark.parametrize("option", ["combination", "permutation"])
def test_device_batch_assembly_matches_host(option):
    """geossl_super_edges vs the itertools-order host restatement of AtomTupleExtractor + collate offsets."""
    from geossl_b200.data import assemble_batch_device, super_edges_host
    counts = [1, 2, 30, 5, 1, 17, 60, 3]
    n = sum(counts)
    z = torch.randint(0, 9, (n,))
    pos = torch.rand(n, 3)
    b = assemble_batch_device(counts, z, pos, option=option, device=DEV)
    assert torch.equal(b.super_edge_index.cpu(), super_edges_host(counts, option))
    assert torch.equal(b.batch.cpu(), torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts)))
    assert b.num_graphs == len(counts) and b.super_edge_index.dtype == torch.int64


@pytest.mark.parametrize("option", ["combination", "permutation"])
def test_device_pair_subsampling(option):
    """--distance_sample_ratio < 1 on the device: int(M * ratio) distinct intra-molecule pairs per molecule; with the
    host draw injected the columns equal the host restatement (= the reference extractor, tests/test_host_logic.py)."""
    import numpy as np
    from geossl_b200.data import assemble_batch_device, pair_count, sample_pairs_host, sampled_pair_count, super_edges_host
    counts, ratio = [1, 2, 30, 5, 1, 17, 60, 3], 0.3
    n = sum(counts)
    z, pos = torch.randint(0, 9, (n,)), torch.rand(n, 3)
    full = super_edges_host(counts, option)
    kept = sampled_pair_count(counts, option, ratio)
    gen = torch.Generator(device=DEV).manual_seed(3)
    b = assemble_batch_device(counts, z, pos, option=option, device=DEV, ratio=ratio, generator=gen)
    sei = b.super_edge_index.cpu()
    assert sei.shape == (2, int(kept.sum()))
    gid = b.batch.cpu()[sei[0]]
    assert torch.equal(gid, b.batch.cpu()[sei[1]])                                   # intra-molecule
    assert torch.equal(torch.bincount(gid, minlength=len(counts)), torch.from_numpy(kept))
    code = lambda e: e[0] * n + e[1]
    assert torch.unique(code(sei)).numel() == sei.shape[1]                           # without replacement
    assert bool(torch.isin(code(sei), code(full)).all())                             # valid pairs of the option
    gen2 = torch.Generator(device=DEV).manual_seed(4)
    b2 = assemble_batch_device(counts, z, pos, option=option, device=DEV, ratio=ratio, generator=gen2)
    assert not torch.equal(b2.super_edge_index, b.super_edge_index)
    np.random.seed(11)
    sel = sample_pairs_host(counts, option, ratio)
    b3 = assemble_batch_device(counts, z, pos, option=option, device=DEV, ratio=ratio, selection=sel)
    assert torch.equal(b3.super_edge_index.cpu(), super_edges_host(counts, option, ratio=ratio, selection=sel))



@pytest.mark.parametrize("owner_small", [True, False])
@pytest.mark.parametrize("ng,lo,hi,density", [(7, 3, 30, 0.05), (4, 40, 70, 0.1), (6, 1, 2, 0.05)])
def test_pair_index_matches_oracle(ng, lo, hi, density, owner_small, monkeypatch):
    """geossl_pair_index vs oracle/radius.py::pair_index, including truncated rows (edges without a reverse)."""
    from oracle.radius import pair_index, radius_neighbors
    monkeypatch.setattr(ops, "PAIR_OWNER_SMALL", owner_small)
    b = synthetic_batch(ng, lo, hi, seed=ng, with_pairs=False, density=density)
    g = ops.radius_csr(b.positions.to(DEV), b.batch.to(DEV), 10.0, num_graphs=ng).ensure_pairs()
    rp, src = radius_neighbors(b.positions, 10.0, b.batch)
    prp, poe, e1, e2 = pair_index(rp, src, owner_small)
    e, u = src.size, len(e1)
    assert g.num_edges == e and int(g.n_pairs_dev.item()) == u
    assert torch.equal(g.pair_rowptr.cpu().long(), torch.from_numpy(prp))
    assert torch.equal(g.pair_of_edge[:e].cpu().long(), torch.from_numpy(poe))
    assert torch.equal(g.pair_e1[:u].cpu().long(), torch.from_numpy(e1))
    assert torch.equal(g.pair_e2[:u].cpu().long(), torch.from_numpy(e2))
    assert torch.equal(g.pair_dist[:u], g.dist[:e][g.pair_e1[:u].long()])
    pa = g.pair_atoms[:u].cpu().long()                                            # (s, t) or (s, ~t) of the canonical edge
    tgt = torch.repeat_interleave(torch.arange(rp.size - 1), torch.from_numpy(np.diff(rp)))
    s_ref, t_ref = torch.from_numpy(src)[e1], tgt[e1]
    both = torch.from_numpy(e2) >= 0
    assert torch.equal(pa[:, 0], s_ref) and torch.equal(torch.where(both, pa[:, 1], ~pa[:, 1]), t_ref)
    assert torch.equal(pa[:, 1] >= 0, both)
    if e:
        rev = g.pair_e2[:u].long()
        has = rev >= 0
        assert torch.equal(g.dist[:e][rev[has]], g.pair_dist[:u][has])          # both directions: the same length, bitwise
    if hi >= 40:
        assert (e2 < 0).any() and u > e // 2                                     # truncation produced orphans


def _same_csr(a, b):
    e = a.num_edges
    assert b.num_edges == e
    assert torch.equal(a.rowptr, b.rowptr) and torch.equal(a.src[:e], b.src[:e]) and torch.equal(a.tgt[:e], b.tgt[:e])
    assert torch.equal(a.dist[:e], b.dist[:e])                       # same fp32 expression -> same bits


@pytest.mark.parametrize("seed,ng,lo,hi,r,density", [
    (0, 7, 1, 9, 10.0, 0.05), (1, 16, 10, 60, 10.0, 0.05), (2, 5, 40, 90, 10.0, 0.08), (3, 6, 25, 70, 3.0, 0.05),
    (4, 3, 200, 600, 6.0, 0.05), (7, 2, 560, 640, 10.0, 0.05), (8, 2, 3000, 5000, 5.0, 0.05), (9, 1, 900, 900, 2.5, 0.4)])
def test_cell_list_is_bit_identical_to_the_index_order_scan(seed, ng, lo, hi, r, density):
    """The spatial cell list (27 cells around the query, index-order select of the 33 smallest in-range atoms) against the
    index-order scan and the CPU oracle: same rowptr / sources / distances, including rows that truncate at 32/33
    neighbours (dense 600-atom pockets at 10 A), rows too dense for the per-warp hit list (density 0.4: scan fall-back),
    tiny graphs, and 5000-atom graphs."""
    b = synthetic_batch(ng, lo, hi, seed=seed, density=density, with_pairs=False)
    pos, bt = b.positions.to(DEV), b.batch.to(DEV)
    g_scan = ops.radius_csr(pos, bt, r, num_graphs=ng, cell_list=False)
    g_cell = ops.radius_csr(pos, bt, r, num_graphs=ng, cell_list=True)
    _same_csr(g_scan, g_cell)
    if b.positions.size(0) <= 2500:
        rowptr, src = radius_neighbors(b.positions, r, b.batch)
        e = int(rowptr[-1])
        assert g_cell.num_edges == e and np.array_equal(g_cell.src[:e].cpu().numpy(), src.astype(np.int32))


def test_cell_list_automatic_threshold_mixes_paths_in_one_batch():
    old = ops.CELL_LIST_MIN_ATOMS
    try:
        ops.CELL_LIST_MIN_ATOMS = 300
        b = synthetic_batch(6, 100, 900, seed=21, with_pairs=False)        # graphs on both sides of the threshold
        pos, bt = b.positions.to(DEV), b.batch.to(DEV)
        _same_csr(ops.radius_csr(pos, bt, 6.0, num_graphs=6, cell_list=False), ops.radius_csr(pos, bt, 6.0, num_graphs=6))
    finally:
        ops.CELL_LIST_MIN_ATOMS = old


def test_fma_contracted_distance_switch():
    """torch_cluster's binary is not available, so both roundings of `dist += (x-y)*(x-y)` are built: per-operation rounding
    (default) and FMA contraction (ops.RADIUS_FMA).  Enumerate atom pairs whose squared distance sits within an ulp of
    r*r: the two variants must disagree on some of them, and each must agree with the oracle evaluated the same way."""
    from oracle.radius import dist2
    rng = np.random.default_rng(0)
    r = np.float32(5.0)
    n_pairs = 60000
    a = (rng.random((n_pairs, 3)) * 20).astype(np.float32)
    u = rng.normal(size=(n_pairs, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    bpts = (a.astype(np.float64) + u * float(r) * (1 + rng.uniform(-2e-7, 2e-7, size=(n_pairs, 1)))).astype(np.float32)
    d = bpts - a
    in_rn, in_fma = dist2(d, False) < r * r, dist2(d, True) < r * r
    differ = np.nonzero(in_rn != in_fma)[0]
    assert differ.size > 10, "the enumeration must reach pairs where the two roundings decide differently"
    sel = np.concatenate([differ, np.arange(200)])
    pos = torch.from_numpy(np.stack([a[sel], bpts[sel]], axis=1).reshape(-1, 3))       # graph k = the two atoms of pair k
    batch = torch.arange(sel.size).repeat_interleave(2)
    old = ops.RADIUS_FMA
    try:
        for fma, expect in ((False, in_rn[sel]), (True, in_fma[sel])):
            ops.RADIUS_FMA = fma
            g = ops.radius_csr(pos.to(DEV), batch.to(DEV), float(r), num_graphs=sel.size, transpose=False)
            deg = (g.rowptr[1:] - g.rowptr[:-1]).cpu().numpy().reshape(-1, 2)
            assert np.array_equal(deg[:, 0] == 1, expect) and np.array_equal(deg[:, 1] == 1, expect)
            assert torch.equal(g.edge_index.cpu(), oracle_radius_graph(pos, float(r), batch, fma=fma))
    finally:
        ops.RADIUS_FMA = old
