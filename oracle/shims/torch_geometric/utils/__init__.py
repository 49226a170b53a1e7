"""Shim for the two ``torch_geometric.utils`` names Geom3D/datasets/datasets_3D.py:7 imports (third-party, un-vendored:
parity unpinned; semantics restated from the PyG 2.0.x sources).  TEST INFRASTRUCTURE."""
import networkx as nx
import torch


def to_networkx(data):
    """Directed graph with nodes 0..num_nodes-1 and one edge per column of ``data.edge_index``."""
    G = nx.DiGraph()
    G.add_nodes_from(range(int(data.x.size(0))))
    for u, v in data.edge_index.t().tolist():
        G.add_edge(u, v)
    return G


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
    """Edges with BOTH endpoints in ``subset`` (original order kept), optionally relabelled to 0..len(subset)-1."""
    if isinstance(subset, (list, tuple)):
        subset = torch.tensor(subset, dtype=torch.long)
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    n_mask = torch.zeros(num_nodes, dtype=torch.bool)
    n_mask[subset] = True
    mask = n_mask[edge_index[0]] & n_mask[edge_index[1]]
    edge_index = edge_index[:, mask]
    edge_attr = edge_attr[mask] if edge_attr is not None else None
    if relabel_nodes:
        n_idx = torch.zeros(num_nodes, dtype=torch.long)
        n_idx[subset] = torch.arange(subset.size(0))
        edge_index = n_idx[edge_index]
    return edge_index, edge_attr
