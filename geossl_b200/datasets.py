"""Collated conformer stores and dataset-time radius graphs.

The reference keeps every processed dataset as one ``geometric_data_processed.pt`` = ``torch.save((data, slices))``:
``data`` is a PyG ``Data`` whose tensors are the per-molecule tensors concatenated, ``slices[key]`` the (M+1,) offsets
(/root/reference/Geom3D/datasets/datasets_3D.py:21,69-80).  ``MoleculeDataset3DRadius.process``
(datasets_3D_Radius.py:105-133; the MD17/LBA/LEP variants do the same) then loops over molecules calling
``radius_graph(data.positions, r=radius, loop=False)`` on the CPU and re-collates.

Here the same layout is a plain ``(dict, dict)`` pair, and the radius graphs of ALL molecules come from one batched
launch sequence of the CSR neighbour kernel (graphs never interact, so the per-molecule results are identical).
"""
import io
import pickle
from itertools import repeat

import numpy as np
import torch

from . import ops
from .data import assemble_batch_device

# keys concatenated along the LAST dimension and holding molecule-local atom indices
# (torch_geometric Data.__cat_dim__ / dataloaders_AtomTuple.py:62-63)
INDEX_KEYS = ("edge_index", "radius_edge_index", "super_edge_index", "full_edge_index")


def _cat_dim(key, item):
    return -1 if (key in INDEX_KEYS or "index" in key) and item.dim() == 2 else 0


class CollatedStore:
    """``(data, slices)`` of an InMemoryDataset: ``get(i)`` slices molecule ``i`` out exactly like
    ``Molecule3DDataset.get`` (datasets_3D.py:69-76)."""

    def __init__(self, data, slices):
        self.data = dict(data)
        self.slices = {k: torch.as_tensor(v, dtype=torch.long) for k, v in slices.items()}
        # keys whose per-molecule edge lists are known to be (target, source)-sorted because THIS package's neighbour
        # kernel wrote them (add_radius_edges).  Lists read from a reference file come from torch_cluster's CPU KD-tree
        # search, whose order inside a target row is not guaranteed: they take the one-time check + stable sort.
        self.sorted_edge_keys = set()

    def __len__(self):
        return int(next(iter(self.slices.values())).numel()) - 1

    def keys(self):
        return list(self.data.keys())

    def get(self, idx):
        out = {}
        for key, item in self.data.items():
            if key not in self.slices:
                continue
            sl = self.slices[key]
            s = list(repeat(slice(None), item.dim()))
            s[_cat_dim(key, item)] = slice(int(sl[idx]), int(sl[idx + 1]))
            out[key] = item[tuple(s)]
        return out

    def atom_counts(self):
        key = "positions" if "positions" in self.slices else "x"
        return (self.slices[key][1:] - self.slices[key][:-1]).numpy()

    # ---- construction / IO
    @staticmethod
    def from_data_list(data_list):
        """The dataset-side ``collate`` (no index increment: indices stay molecule-local)."""
        keys = list(data_list[0].keys())
        data, slices = {}, {}
        for k in keys:
            items = [torch.as_tensor(d[k]) for d in data_list]
            items = [t.unsqueeze(0) if t.dim() == 0 else t for t in items]
            dim = _cat_dim(k, items[0])
            data[k] = torch.cat(items, dim=dim)
            slices[k] = torch.tensor([0] + list(np.cumsum([t.size(dim) for t in items])), dtype=torch.long)
        return CollatedStore(data, slices)

    def save(self, path):
        torch.save((self.data, {k: v for k, v in self.slices.items()}), path)

    @staticmethod
    def load(path):
        """Reads a store written by ``save`` or a reference ``geometric_data_processed.pt`` (PyG ``Data`` pickled
        inside; torch_geometric itself is not needed -- its classes are mapped onto attribute bags while unpickling)."""
        obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)
        data, slices = obj
        return CollatedStore(_as_tensor_dict(data), _as_tensor_dict(slices))


class _Bag:
    """Stand-in for torch_geometric's Data / storage classes while unpickling."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("torch_geometric"):
            return type(name, (_Bag,), {})
        return super().find_class(module, name)


class _TolerantPickle:
    """``pickle_module`` for torch.load: everything from ``torch_geometric.*`` becomes a ``_Bag``."""
    __name__ = "pickle"
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _Unpickler(io.BytesIO(b), **kw).load())
    dump, dumps, Pickler = pickle.dump, pickle.dumps, pickle.Pickler
    HIGHEST_PROTOCOL, DEFAULT_PROTOCOL = pickle.HIGHEST_PROTOCOL, pickle.DEFAULT_PROTOCOL
    UnpicklingError, PicklingError = pickle.UnpicklingError, pickle.PicklingError


def _as_tensor_dict(obj):
    """dict | PyG<2 Data (tensors in __dict__) | PyG>=2 Data (tensors in _store._mapping) -> {key: tensor}."""
    if isinstance(obj, dict):
        d = obj
    else:
        d = dict(getattr(obj, "__dict__", {}))
        store = d.pop("_store", None)
        if store is not None:
            d = dict(getattr(store, "__dict__", {}).get("_mapping", {}))
    return {k: v for k, v in d.items() if torch.is_tensor(v)}


def add_radius_edges(store, radius, device="cuda", max_num_neighbors=32, chunk_atoms=4_000_000, key="radius_edge_index"):
    """``MoleculeDataset3DRadius.process`` (datasets_3D_Radius.py:117-121) for the whole store at once.

    Molecules are concatenated into chunks of at most ``chunk_atoms`` atoms with a graph-id vector; each chunk is ONE
    neighbour-search launch.  The (source, target) lists come back target-sorted per molecule; they are shifted to
    molecule-local indices and sliced per molecule from the row offsets -- the layout the reference's per-molecule loop
    + collate produces."""
    counts = store.atom_counts()
    pos_all = store.data["positions"].to(torch.float32)
    atom_off = np.concatenate([[0], np.cumsum(counts)])
    m = len(counts)
    pieces, per_mol = [], []
    lo = 0
    while lo < m:
        hi = lo + 1
        while hi < m and atom_off[hi + 1] - atom_off[lo] <= chunk_atoms:
            hi += 1
        a0, a1 = int(atom_off[lo]), int(atom_off[hi])
        bvec = torch.from_numpy(np.repeat(np.arange(hi - lo, dtype=np.int64), counts[lo:hi])).to(device)
        g = ops.radius_csr(pos_all[a0:a1].to(device), bvec, radius, max_num_neighbors, num_graphs=hi - lo, transpose=False)
        e = g.num_edges                                           # one host sync per chunk (dataset-time, not per step)
        src, tgt = g.src[:e].long(), g.tgt[:e].long()
        local = torch.from_numpy(atom_off[lo:hi] - a0).to(device)[bvec]          # first atom of each atom's molecule
        pieces.append(torch.stack([src - local[src], tgt - local[tgt]]).cpu())
        rp = g.rowptr.long().cpu().numpy()
        per_mol.append(rp[atom_off[lo + 1:hi + 1] - a0] - rp[atom_off[lo:hi] - a0])
        lo = hi
    store.data[key] = torch.cat(pieces, dim=1) if pieces else torch.empty((2, 0), dtype=torch.long)
    store.sorted_edge_keys.add(key)
    store.slices[key] = torch.from_numpy(np.concatenate([[0], np.cumsum(np.concatenate(per_mol))]) if per_mol
                                         else np.zeros(1, np.int64)).long()
    return store


def batch_from_store(store, indices, device="cuda", option="combination", ratio=1.0, generator=None, mask_ratio=0.0,
                     rng=np.random):
    """One training batch (``DataLoaderAtomTuple`` + ``AtomTupleExtractor``, dataloaders_AtomTuple.py:15-73,81-90) for
    the molecules ``indices`` of a store, assembled on the device: ``batch``, ``super_edge_index`` by kernel,
    ``radius_edge_index`` (if the store has it) offset by the cumulative atom count.  ``mask_ratio > 0`` applies the
    atom-masking augmentation to every molecule first, like ``Molecule3DDataset.get`` (datasets_3D.py:77-78)."""
    mols = [store.get(int(i)) for i in indices]
    if mask_ratio > 0:
        mols = [mask_subgraph(d, mask_ratio, rng) for d in mols]
    counts = [int(d["positions"].size(0)) for d in mols]
    z = torch.cat([d["x"][:, 0] if d["x"].dim() == 2 else d["x"] for d in mols]).long()
    pos = torch.cat([d["positions"] for d in mols]).to(torch.float32)
    b = assemble_batch_device(counts, z, pos, option=option, device=device, ratio=ratio, generator=generator)
    if "radius_edge_index" in store.data:
        off = np.concatenate([[0], np.cumsum(counts)])
        b.radius_edge_index = torch.cat([d["radius_edge_index"] + int(off[i]) for i, d in enumerate(mols)], dim=1).to(device)
        # only lists written by add_radius_edges are known to be sorted (per-molecule target-sorted, molecule-major)
        b.extras["rei_sorted"] = "radius_edge_index" in store.sorted_edge_keys
    return b


def mask_subgraph(mol, mask_ratio, rng=np.random):
    """The atom-masking augmentation of ``Molecule3DDataset.subgraph`` (datasets_3D.py:24-67,77-78) on one molecule
    ``mol`` = the dict ``CollatedStore.get`` returns: a random connected-ish sub-molecule is grown over the bond graph
    (``edge_index``) and everything else is dropped.

    Host logic, faithful to the reference including its draws and quirks, so that with ``rng=np.random`` and the same
    ``np.random.seed`` the kept atoms are the reference's: the start atom is ``randint(node_num, size=1)[0]`` (:29); each
    step draws ``choice(list(frontier))`` from a Python *set* of not-yet-kept successors (:37-42, set iteration order
    included), re-seeding the frontier with ``choice`` over the sorted unvisited atoms when it runs dry (:34-36); the loop
    runs ``while len(kept) <= int(n * (1 - mask_ratio))`` and therefore keeps one atom more than that count (:27,33).
    ``edge_index`` / ``edge_attr`` / ``radius_edge_index`` keep the edges with both endpoints kept, in their original
    order, relabelled to the compacted atom numbering (PyG ``subgraph(..., relabel_nodes=True)``, :47-65); ``x`` and
    ``positions`` are sliced.  Returns a new dict (other keys are passed through)."""
    x = mol["x"]
    node_num = int(x.size(0))
    sub_num = int(node_num * (1 - mask_ratio))
    ei = mol["edge_index"]
    succ = [[] for _ in range(node_num)]
    for u, v in ei.t().tolist():                                  # successor lists of the directed bond graph (:25)
        succ[u].append(v)
    idx_sub = [rng.randint(node_num, size=1)[0]]
    idx_neigh = set([n for n in succ[idx_sub[-1]]])
    while len(idx_sub) <= sub_num:
        if len(idx_neigh) == 0:
            idx_unsub = list(set([n for n in range(node_num)]).difference(set(idx_sub)))
            idx_neigh = set([rng.choice(idx_unsub)])
        sample_node = rng.choice(list(idx_neigh))
        idx_sub.append(sample_node)
        idx_neigh = idx_neigh.union(set([n for n in succ[idx_sub[-1]]])).difference(set(idx_sub))
    keep = sorted(int(i) for i in idx_sub)
    keep_t = torch.tensor(keep, dtype=torch.long)
    n_mask = torch.zeros(node_num, dtype=torch.bool)
    n_mask[keep_t] = True
    n_idx = torch.zeros(node_num, dtype=torch.long)
    n_idx[keep_t] = torch.arange(len(keep))

    def sub_edges(edge_index, edge_attr=None):
        m = n_mask[edge_index[0]] & n_mask[edge_index[1]]
        return n_idx[edge_index[:, m]], (None if edge_attr is None else edge_attr[m])

    out = dict(mol)
    out["edge_index"], ea = sub_edges(ei, mol.get("edge_attr"))
    if ea is not None:
        out["edge_attr"] = ea
    out["x"] = x[keep_t]
    out["positions"] = mol["positions"][keep_t]
    if "radius_edge_index" in mol:
        out["radius_edge_index"], _ = sub_edges(mol["radius_edge_index"])
    return out
