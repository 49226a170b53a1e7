"""Shim for the two PyG names schnet.py:12 imports.  Semantics restated from SURVEY.md
Appendix B.1 / B.3 (third-party, parity unpinned)."""
import torch

from oracle.radius import radius_graph  # noqa: F401  (same signature as PyG's)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        assert aggr == "add"
        self.aggr = aggr

    def propagate(self, edge_index, **kwargs):
        x = kwargs.pop("x")
        x_j = x.index_select(0, edge_index[0])
        msg = self.message(x_j, **kwargs)
        out = torch.zeros((x.size(0),) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add(0, edge_index[1], msg)
