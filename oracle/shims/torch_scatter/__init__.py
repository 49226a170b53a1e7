"""Shim for torch_scatter.scatter / scatter_add (parity unpinned)."""
import torch


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    if dim < 0:
        dim = src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add(dim, index, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add(
            0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1)
        view = [1] * src.dim()
        view[dim] = dim_size
        return res / cnt.view(view)
    raise ValueError(reduce)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")
