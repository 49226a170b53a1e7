"""CPU restatement of ``torch_cluster.radius_graph`` as the reference calls it
(/root/reference/Geom3D/models/schnet.py:91, datasets_3D_Radius.py:120).  TEST INFRASTRUCTURE.

torch_cluster is an un-vendored, un-pinned third-party dependency (README.md:23-24 pins only
pytorch=1.9.1 / pyg=2.0.2; contemporaneous torch-cluster 1.5.9) and is absent from this image:
**parity unpinned**.  The semantics restated here are the published CUDA ``radius_kernel``
(SURVEY.md Appendix B.1):

  for every query atom y, scan the atoms x of the same graph in ascending index order;
  dist = ((dx*dx) + dy*dy) + dz*dz in fp32, no FMA contraction; keep x if dist < r*r (r*r in fp32);
  stop after max_num_neighbors+1 = 33 hits (self included); emit (source=x, target=y);
  drop self pairs.  Output (2,E) int64, grouped by target ascending, sources ascending.

numpy float32 element-wise ops round once per op, so the boundary decision is bit-identical to a
CUDA kernel using __fsub_rn/__fmul_rn/__fadd_rn.
"""
import numpy as np
import torch


def _graph_ptr(batch_np, num_graphs=None):
    if batch_np.size == 0:
        return np.zeros(1, dtype=np.int64)
    assert np.all(batch_np[1:] >= batch_np[:-1]), "batch must be sorted (torch_cluster requirement)"
    nb = int(batch_np[-1]) + 1 if num_graphs is None else int(num_graphs)
    counts = np.bincount(batch_np, minlength=nb)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


def dist2(d, fma=False):
    """Squared length of the fp32 difference vectors ``d`` (...,3).  Default: ((dx*dx) + dy*dy) + dz*dz, one fp32 rounding
    per operation.  ``fma=True``: fma(dz, dz, fma(dy, dy, dx*dx)) -- what `dist += (x-y)*(x-y)` becomes when the compiler
    contracts it (nvcc -fmad=true).  The fused steps are emulated in float64 (a 24 x 24-bit product is exact there; the
    following addition can, very rarely, round twice -- acceptable for test data, noted here)."""
    if not fma:
        sq = d * d
        return (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    d64 = d.astype(np.float64)
    t = (d[..., 0] * d[..., 0]).astype(np.float32)                              # dx*dx rounded to fp32
    t = (d64[..., 1] * d64[..., 1] + t.astype(np.float64)).astype(np.float32)   # fma(dy, dy, t)
    return (d64[..., 2] * d64[..., 2] + t.astype(np.float64)).astype(np.float32)


def radius_neighbors(pos, r, batch=None, max_num_neighbors=32, fma=False):
    """Returns (rowptr int64 (N+1,), src int64 (E,)) -- the destination-sorted CSR."""
    p = np.ascontiguousarray(pos.detach().cpu().numpy() if torch.is_tensor(pos) else pos, dtype=np.float32)
    n_atoms = p.shape[0]
    if batch is None:
        b = np.zeros(n_atoms, dtype=np.int64)
    else:
        b = batch.detach().cpu().numpy().astype(np.int64) if torch.is_tensor(batch) else np.asarray(batch, np.int64)
    ptr = _graph_ptr(b)
    r2 = np.float32(r) * np.float32(r)
    limit = max_num_neighbors + 1
    deg = np.zeros(n_atoms, dtype=np.int64)
    rows = []
    for g in range(len(ptr) - 1):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        if hi <= lo:
            continue
        x = p[lo:hi]                                   # candidates (ascending index)
        d = x[None, :, :] - x[:, None, :]              # d[y, x] = pos[x] - pos[y]
        dist = dist2(d, fma)                           # fp32, see dist2
        hit = dist < r2
        rank = np.cumsum(hit, axis=1)                  # 1-based rank among hits
        keep = hit & (rank <= limit)
        keep[np.arange(hi - lo), np.arange(hi - lo)] = False
        for y in range(hi - lo):
            nb = np.nonzero(keep[y])[0] + lo
            deg[lo + y] = nb.size
            rows.append(nb)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    src = np.concatenate(rows).astype(np.int64) if rows else np.zeros(0, dtype=np.int64)
    return rowptr, src


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", fma=False):
    """Same signature as ``torch_geometric.nn.radius_graph``; (2,E) int64 ``[source; target]``."""
    assert not loop and flow == "source_to_target"
    rowptr, src = radius_neighbors(x, r, batch, max_num_neighbors, fma)
    deg = np.diff(rowptr)
    tgt = np.repeat(np.arange(deg.size, dtype=np.int64), deg)
    ei = torch.from_numpy(np.stack([src, tgt]).astype(np.int64))
    return ei.to(x.device) if torch.is_tensor(x) else ei


def transpose_csr(rowptr, src):
    """Source-sorted view of the same edge list: (t_rowptr, t_eid, t_tgt); edges of one source are
    ordered by ascending target (= ascending edge id, a stable counting sort)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    src = np.asarray(src, dtype=np.int64)
    n = rowptr.size - 1
    tgt = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    order = np.argsort(src, kind="stable")
    t_rowptr = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=n))]).astype(np.int64)
    return t_rowptr, order.astype(np.int64), tgt[order]


def pair_index(rowptr, src, owner_small=True):
    """Undirected-pair index of a destination-sorted CSR with ascending sources per row (the contract of
    geossl_pair_index, include/geossl_b200.h): pair ids are assigned row by row to the canonical edges -- source >
    target (``owner_small``) or source < target, or an edge whose reverse is absent -- in row order.  Returns (pair_rowptr, pair_of_edge, pair_e1, pair_e2)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    src = np.asarray(src, dtype=np.int64)
    n = rowptr.size - 1
    tgt = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    eid = {(int(s), int(t)): e for e, (s, t) in enumerate(zip(src, tgt))}
    pair_of_edge = np.full(src.size, -1, dtype=np.int64)
    pair_rowptr = np.zeros(n + 1, dtype=np.int64)
    e1, e2 = [], []
    for e, (s, t) in enumerate(zip(src.tolist(), tgt.tolist())):
        rev = eid.get((t, s), -1)
        if (s > t if owner_small else s < t) or rev < 0:
            pair_of_edge[e] = len(e1)
            e1.append(e)
            e2.append(rev)
            pair_rowptr[t + 1] += 1
    for e, (s, t) in enumerate(zip(src.tolist(), tgt.tolist())):
        if pair_of_edge[e] < 0:
            pair_of_edge[e] = pair_of_edge[eid[(t, s)]]
    return np.cumsum(pair_rowptr), pair_of_edge, np.asarray(e1, dtype=np.int64), np.asarray(e2, dtype=np.int64)
