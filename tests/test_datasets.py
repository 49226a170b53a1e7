"""Collated stores (geometric_data_processed.pt layout) and dataset-time radius graphs."""
import sys
import types

import numpy as np
import pytest
import torch

from geossl_b200.datasets import CollatedStore
from oracle.radius import radius_graph


def _molecules(counts, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in counts:
        ei = torch.randint(0, max(n, 1), (2, 2 * n), generator=g)
        out.append(dict(x=torch.randint(0, 9, (n, 2), generator=g), positions=torch.rand(n, 3, generator=g) * 6.0,
                        edge_index=ei, y=torch.randn((), generator=g)))
    return out


def test_store_roundtrip_and_get(tmp_path):
    mols = _molecules([3, 1, 7, 4])
    st = CollatedStore.from_data_list(mols)
    assert len(st) == 4 and st.atom_counts().tolist() == [3, 1, 7, 4]
    assert st.data["edge_index"].shape == (2, 30)                       # indices concatenated along the last dim
    p = tmp_path / "geometric_data_processed.pt"
    st.save(p)
    st2 = CollatedStore.load(p)
    for i, m in enumerate(mols):
        d = st2.get(i)
        assert torch.equal(d["x"], m["x"]) and torch.equal(d["positions"], m["positions"])
        assert torch.equal(d["edge_index"], m["edge_index"]) and torch.equal(d["y"], m["y"].view(1))


@pytest.mark.parametrize("layout", ["dict_attrs", "store_mapping"])
def test_load_reference_file_without_torch_geometric(tmp_path, layout):
    """A (Data, slices) pickle written by a PyG-like class must load when torch_geometric is not importable."""
    mols = _molecules([2, 5, 3], seed=1)
    ref = CollatedStore.from_data_list(mols)
    names = ["torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.data.storage"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        mods = {n: types.ModuleType(n) for n in names}
        Data = type("Data", (), {"__module__": "torch_geometric.data.data"})
        GlobalStorage = type("GlobalStorage", (), {"__module__": "torch_geometric.data.storage"})
        mods["torch_geometric.data.data"].Data = Data
        mods["torch_geometric.data.storage"].GlobalStorage = GlobalStorage
        sys.modules.update(mods)
        data = Data()
        if layout == "dict_attrs":                                    # PyG < 2: tensors are plain attributes
            data.__dict__.update(ref.data)
        else:                                                          # PyG >= 2: Data._store._mapping
            st = GlobalStorage()
            st.__dict__["_mapping"] = dict(ref.data)
            data.__dict__["_store"] = st
        p = tmp_path / "geometric_data_processed.pt"
        torch.save((data, dict(ref.slices)), p)
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    assert "torch_geometric.data.data" not in sys.modules or saved["torch_geometric.data.data"] is not None
    got = CollatedStore.load(p)
    assert sorted(got.keys()) == sorted(ref.keys())
    for i in range(3):
        a, b = got.get(i), ref.get(i)
        assert all(torch.equal(a[k], b[k]) for k in b)


@pytest.mark.gpu
def test_dataset_time_radius_edges_match_per_molecule_oracle():
    from geossl_b200.datasets import add_radius_edges, batch_from_store
    counts = [1, 9, 40, 2, 17, 60, 5]
    st = CollatedStore.from_data_list(_molecules(counts, seed=2))
    add_radius_edges(st, 3.0, device="cuda:0", chunk_atoms=64)       # several chunks
    assert st.slices["radius_edge_index"].numel() == len(counts) + 1
    for i in range(len(counts)):
        d = st.get(i)
        want = radius_graph(d["positions"], 3.0, None)               # datasets_3D_Radius.py:120, one molecule at a time
        assert torch.equal(d["radius_edge_index"], want), i
    b = batch_from_store(st, [5, 2, 0, 4], device="cuda:0")
    assert b.num_graphs == 4 and b.positions.shape[0] == 60 + 40 + 1 + 17
    want = radius_graph(b.positions.cpu(), 3.0, b.batch.cpu())
    assert torch.equal(b.radius_edge_index.cpu(), want)
    assert bool((b.batch[b.super_edge_index[0]] == b.batch[b.super_edge_index[1]]).all())


def test_mask_subgraph_matches_reference_fixture():
    """The atom-masking augmentation (datasets_3D.py:24-67) replays the reference's numpy draws: same kept atoms, same
    relabelled bond / radius edges, for every molecule of the fixture written by the UNMODIFIED reference function
    (tests/golden/make_golden.py::masking_case), including a molecule with a disconnected fragment."""
    from _golden import Golden
    from geossl_b200.datasets import mask_subgraph
    g = Golden("masking_small")
    c, i, o = g.cfg, g["in"], g["out"]
    np.random.seed(c["seed"])
    for m in range(c["n_mol"]):
        mol = {k: i[f"{k}{m}"] for k in ("x", "positions", "edge_index", "edge_attr", "radius_edge_index")}
        got = mask_subgraph(mol, c["mask_ratio"])
        n = mol["x"].size(0)
        assert got["x"].size(0) == int(n * (1 - c["mask_ratio"])) + 1           # the reference's off-by-one (:33)
        for k in ("x", "positions", "edge_index", "edge_attr", "radius_edge_index"):
            assert torch.equal(got[k], o[f"{k}{m}"]), (m, k)
        assert torch.equal(mol["x"], i[f"x{m}"])                                # input left untouched


def test_mask_subgraph_keeps_edges_consistent():
    from geossl_b200.datasets import mask_subgraph
    rng = np.random.RandomState(3)
    mol = _molecules([20], seed=5)[0]
    out = mask_subgraph(mol, 0.5, rng)
    k = out["x"].size(0)
    assert k == 11 and out["positions"].shape == (k, 3)
    assert out["edge_index"].numel() == 0 or int(out["edge_index"].max()) < k
