"""Synthetic Molecule3D / MD17 / LBA-shaped conformer batches and the AtomTuple batch contract.

Host-side mirror of the input layout the hot path consumes
(/root/reference/Geom3D/dataloaders/dataloaders_AtomTuple.py:15-37 ``AtomTupleExtractor`` and
:45-78 ``BatchAtomTuple.from_data_list``): ``x`` (N,2) int64 with the atom class in column 0
(datasets_utils.py:131), ``positions`` (N,3) fp32, ``batch`` (N,) int64 sorted, and
``super_edge_index`` (2,P) int64 -- all ordered atom pairs of each molecule, graph-major, either
``combination`` (i<j, n(n-1)/2) or ``permutation`` (i!=j, n(n-1)), offset by the cumulative node
count.  Data is synthetic (no datasets in this image): 9 atom classes (pretrain_GeoSSL.py:309),
positions uniform in a cube at ~0.05 atoms/A^3.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch


@dataclass
class AtomTupleBatch:
    """Duck-type of ``BatchAtomTuple``: the attributes do_DDM / NCSN_version_03 read."""
    x: torch.Tensor
    positions: torch.Tensor
    batch: torch.Tensor
    super_edge_index: torch.Tensor
    radius_edge_index: Optional[torch.Tensor] = None
    n_graphs: Optional[int] = None          # cached; the reference syncs on batch[-1].item()+1
    graph_ptr: Optional[torch.Tensor] = None  # (B+1,) int32 atom offsets (device CSR helper)
    extras: dict = field(default_factory=dict)

    @property
    def num_graphs(self):
        if self.n_graphs is None:
            self.n_graphs = int(self.batch[-1].item()) + 1   # dataloaders_AtomTuple.py:75-78
        return self.n_graphs

    def to(self, device, non_blocking=False):
        def mv(t):
            return None if t is None else t.to(device, non_blocking=non_blocking)
        extras = {k: (mv(v) if torch.is_tensor(v) else v) for k, v in self.extras.items()}
        return AtomTupleBatch(mv(self.x), mv(self.positions), mv(self.batch), mv(self.super_edge_index),
                              mv(self.radius_edge_index), self.n_graphs, mv(self.graph_ptr), extras)

    def pin_memory(self):
        def pm(t):
            return None if t is None else t.pin_memory()
        extras = {k: (pm(v) if torch.is_tensor(v) else v) for k, v in self.extras.items()}
        return AtomTupleBatch(pm(self.x), pm(self.positions), pm(self.batch), pm(self.super_edge_index),
                              pm(self.radius_edge_index), self.n_graphs, pm(self.graph_ptr), extras)


def pair_count(n, option="combination"):
    n = np.asarray(n, dtype=np.int64)
    full = n * (n - 1)
    return full // 2 if option == "combination" else full


def sampled_pair_count(n, option="combination", ratio=1.0):
    """``int(M * ratio)`` pairs kept per molecule (dataloaders_AtomTuple.py:25-27); M itself when ratio >= 1."""
    m = pair_count(n, option)
    if ratio >= 1:
        return m
    return np.asarray([int(v * ratio) for v in np.atleast_1d(m).tolist()], dtype=np.int64).reshape(np.shape(m))


def sample_pairs_host(counts, option="combination", ratio=1.0, rng=np.random):
    """The reference's sub-sampling draw (dataloaders_AtomTuple.py:25-29): per molecule, in molecule order,
    ``rng.choice(M, int(M * ratio), replace=False)``.  With ``rng=np.random`` and the same ``np.random.seed`` this
    consumes the global numpy stream exactly like ``AtomTupleExtractor.__call__``.  Returns one index array per
    molecule (molecules with fewer than two atoms draw nothing, :17)."""
    sel = []
    for n in np.asarray(counts, dtype=np.int64).tolist():
        m = int(pair_count(n, option))
        sel.append(rng.choice(m, int(m * ratio), replace=False).astype(np.int64) if n >= 2 else np.empty(0, np.int64))
    return sel


def super_edges_host(counts, option="combination", ratio=1.0, selection=None, rng=np.random):
    """All ordered atom pairs per molecule in ``itertools`` order, offset by cumulative atom count.
    Vectorised numpy restatement of AtomTupleExtractor + the collate offset; ``ratio < 1`` keeps the sampled
    columns in draw order (``selection`` = the per-molecule index arrays, default: drawn from ``rng``)."""
    counts = np.asarray(counts, dtype=np.int64)
    if ratio < 1 and selection is None:
        selection = sample_pairs_host(counts, option, ratio, rng)
    offs = np.concatenate([[0], np.cumsum(counts)])
    us, vs = [], []
    cache = {}
    for g, n in enumerate(counts.tolist()):
        if n < 2:
            continue
        if n not in cache:
            if option == "combination":
                u, v = np.triu_indices(n, k=1)
            else:
                u = np.repeat(np.arange(n), n - 1)
                v = np.concatenate([np.delete(np.arange(n), i) for i in range(n)])
            cache[n] = (u.astype(np.int64), v.astype(np.int64))
        u, v = cache[n]
        if selection is not None:
            u, v = u[selection[g]], v[selection[g]]
        us.append(u + offs[g])
        vs.append(v + offs[g])
    if not us:
        return torch.empty((2, 0), dtype=torch.long)
    return torch.from_numpy(np.stack([np.concatenate(us), np.concatenate(vs)]))


def synthetic_batch(num_graphs=32, atoms=30, atoms_max=None, *, seed=0, option="combination",
                    density=0.05, node_class=9, with_pairs=True):
    """One synthetic batch on the host.  ``atoms_max`` set => n ~ U{atoms..atoms_max} per graph."""
    rng = np.random.default_rng(seed)
    if atoms_max is None:
        counts = np.full(num_graphs, atoms, dtype=np.int64)
    else:
        counts = rng.integers(atoms, atoms_max + 1, size=num_graphs).astype(np.int64)
    n_total = int(counts.sum())
    z = rng.integers(0, node_class, size=n_total).astype(np.int64)
    side = np.repeat((counts / density) ** (1.0 / 3.0), counts)
    pos = (rng.random((n_total, 3)) * side[:, None]).astype(np.float32)
    batch = np.repeat(np.arange(num_graphs, dtype=np.int64), counts)
    x = np.stack([z, np.zeros_like(z)], axis=1)
    sei = super_edges_host(counts, option) if with_pairs else torch.empty((2, 0), dtype=torch.long)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return AtomTupleBatch(torch.from_numpy(x), torch.from_numpy(pos), torch.from_numpy(batch), sei,
                          None, int(num_graphs), torch.from_numpy(ptr), {"max_graph_atoms": int(counts.max())})


def assemble_batch_device(counts, z, positions, option="combination", device="cuda", ratio=1.0, selection=None,
                          generator=None):
    """Batch assembly on the GPU from per-molecule atom counts (host list) + concatenated ``z`` / ``positions``:
    ``batch``, ``graph_ptr`` and ``super_edge_index`` are produced by geossl_super_edges instead of the reference's
    Python itertools loop (dataloaders_AtomTuple.py:15-37) and collate (:45-73).

    ``ratio < 1`` (``--distance_sample_ratio``, :25-29) keeps ``int(M * ratio)`` distinct pairs per molecule.  The
    draw happens on the device (random key per pair, one sort keyed by (molecule, key), first ``int(M * ratio)`` of
    each segment): same distribution as ``np.random.choice(..., replace=False)``, a different random stream.  Pass
    ``selection`` (per-molecule index arrays, e.g. from ``sample_pairs_host``) to reproduce the host draw exactly."""
    from . import ops
    counts = np.asarray(counts, dtype=np.int64)
    n_atoms, n_graphs = int(counts.sum()), len(counts)
    pc = pair_count(counts, option)
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)).to(device)
    pair_off = np.concatenate([[0], np.cumsum(pc)]).astype(np.int64)
    pair_ptr = torch.from_numpy(pair_off).to(device)
    sei, batch = ops.super_edges(ptr, pair_ptr, n_graphs, n_atoms, int(pc.sum()), option == "permutation")
    if ratio < 1:
        if selection is not None:
            cols = np.concatenate([np.asarray(s, dtype=np.int64) + pair_off[g] for g, s in enumerate(selection)]
                                  + [np.empty(0, np.int64)])
            cols = torch.from_numpy(cols).to(device)
        else:
            kept = sampled_pair_count(counts, option, ratio)                    # host-known => no device sync
            kept_off = np.concatenate([[0], np.cumsum(kept)]).astype(np.int64)
            seg = torch.repeat_interleave(torch.arange(n_graphs, device=device), torch.from_numpy(pc).to(device),
                                          output_size=int(pc.sum()))
            key = seg.double() + torch.rand(seg.numel(), device=device, dtype=torch.float64, generator=generator)
            order = torch.argsort(key)                                          # molecule-major, random within
            seg_k = torch.repeat_interleave(torch.arange(n_graphs, device=device), torch.from_numpy(kept).to(device),
                                            output_size=int(kept.sum()))
            rank = torch.arange(int(kept.sum()), device=device) - torch.from_numpy(kept_off).to(device)[seg_k]
            cols = order[pair_ptr[seg_k] + rank]
        sei = sei[:, cols].contiguous()
    z = z.to(device)
    x = torch.stack([z, torch.zeros_like(z)], dim=1)
    return AtomTupleBatch(x, positions.to(device), batch, sei, None, n_graphs, ptr,
                          {"max_graph_atoms": int(counts.max()) if len(counts) else 0})


# ------------------------------------------------------------------------------------------ capacity padding
PAD_SPACING = 128.0      # Angstrom between padding atoms: larger than any cutoff, so they never gain an edge


def pad_batch(batch, n_atoms_cap, n_pairs_cap, n_edges_cap=None, max_graph_atoms_cap=None):
    """Capacity-padded copy of ``batch`` (same device): fixed tensor shapes for every batch of a stream, which is what
    lets ONE captured CUDA graph serve variable-size Molecule3D batches (10-60 atoms per molecule,
    datasets_utils.py:112-176; the collate is dataloaders_AtomTuple.py:45-78).

    * atoms ``[N, n_atoms_cap)`` are padding: class 0, graph id ``B`` (ONE extra, trailing graph -- ``num_graphs``
      becomes ``B + 1`` for every batch), placed ``PAD_SPACING`` apart on a line far from the molecules, so the
      neighbour search gives them no edge; nothing downstream reads their rows, their gradients are exact zeros.
    * pairs ``[P, n_pairs_cap)`` are padding (index 0/0); ``extras['n_pairs_live']`` = (1,) int32 ``P`` travels with
      the batch and the DDM head kernels read it on the device (geossl_ddm_head_*: n_pairs_live).  The loss is a mean
      over ``max(graph id of a LIVE pair) + 1`` graphs, so the padding graph does not enter it (NCSN.py:210-212).
    * ``radius_edge_index`` (PaiNN), if present, is padded to ``n_edges_cap`` columns ``[idx_i = 0; idx_j = N_cap]``:
      the sentinel ``idx_j`` (one past the last atom) keeps the list sorted and makes ``rowptr[N_cap]`` the live edge
      count on the device (geossl_painn_edge_geometry).  Because two stacked views cannot simply be concatenated any
      more (the sentinel of view 1 is a real atom of view 2), ``extras['rei_stacked']`` (2, 2 n_edges_cap) holds the
      stacked list ready made: live edges of view 1, live edges of view 2 (+ N_cap), padding ``[0; 2 N_cap]``.
    * ``extras['max_graph_atoms']`` (host-known bound on the atoms of a molecule) is replaced by ``max_graph_atoms_cap``
      when given: the bound of the whole stream, since a captured graph bakes the kernel configuration in.
    """
    has_pairs = batch.super_edge_index is not None
    n, p, b = batch.positions.size(0), (batch.super_edge_index.size(1) if has_pairs else 0), batch.num_graphs
    if n > n_atoms_cap or p > n_pairs_cap:
        raise ValueError(f"batch ({n} atoms, {p} pairs) exceeds the capacity ({n_atoms_cap}, {n_pairs_cap})")
    dev = batch.positions.device
    k = n_atoms_cap - n
    x = torch.zeros((n_atoms_cap, batch.x.size(1)), dtype=batch.x.dtype, device=dev)
    x[:n] = batch.x
    pos = torch.zeros((n_atoms_cap, 3), dtype=batch.positions.dtype, device=dev)
    pos[:n] = batch.positions
    if k:
        pos[n:, 0] = 1.0e4 + PAD_SPACING * torch.arange(k, dtype=pos.dtype, device=dev)
    bvec = torch.full((n_atoms_cap,), b, dtype=batch.batch.dtype, device=dev)
    bvec[:n] = batch.batch
    extras = dict(batch.extras)
    sei = None
    if has_pairs:
        sei = torch.zeros((2, n_pairs_cap), dtype=batch.super_edge_index.dtype, device=dev)
        sei[:, :p] = batch.super_edge_index
        extras["n_pairs_live"] = torch.tensor([p], dtype=torch.int32, device=dev)
    if max_graph_atoms_cap is not None:                # one bound for the whole stream (a captured graph bakes it in); the
        extras["max_graph_atoms"] = int(max_graph_atoms_cap)   # edge-less padding graph may exceed it (geossl_cfconv_pairs)
    extras["n_atoms_live"] = n
    extras["n_graphs_live"] = b                     # fine-tune readouts drop the padding graph's row (finetune._readout)
    rei = batch.radius_edge_index
    if rei is not None and n_edges_cap is not None:
        e = rei.size(1)
        if e > n_edges_cap:
            raise ValueError(f"batch has {e} radius edges, capacity is {n_edges_cap}")
        rei_p = torch.zeros((2, n_edges_cap), dtype=rei.dtype, device=dev)
        rei_p[1] = n_atoms_cap
        rei_p[:, :e] = rei
        st = torch.zeros((2, 2 * n_edges_cap), dtype=rei.dtype, device=dev)
        st[1] = 2 * n_atoms_cap
        st[:, :e] = rei
        st[:, e:2 * e] = rei + n_atoms_cap
        extras["rei_stacked"] = st
        rei = rei_p
    ptr = None
    if batch.graph_ptr is not None:
        ptr = torch.cat([batch.graph_ptr.to(torch.int32), torch.tensor([n_atoms_cap], dtype=torch.int32, device=dev)])
    return AtomTupleBatch(x, pos, bvec, sei, rei, b + 1, ptr, extras)
