"""Stand-in for ``torch_scatter`` (2.1.2): ``scatter`` with reduce in {sum, add, mean, max} -- TEST INFRASTRUCTURE,
see ../README.md."""
import torch

__version__ = "2.1.2+segnn_b200_shim"


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    dim = dim % src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    if index.dim() == 1 and src.dim() > 1:
        view = [1] * src.dim()
        view[dim] = -1
        index_b = index.view(view).expand_as(src)
    else:
        index_b = index
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, index_b, src)
    if reduce == "mean":
        total = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, index_b, src)
        count = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add_(
            0, index if index.dim() == 1 else index.select(-1, 0), torch.ones(index.shape[0], dtype=src.dtype,
                                                                                 device=src.device))
        view = [1] * src.dim()
        view[dim] = -1
        return total / count.clamp_(min=1).view(view)
    if reduce == "max":
        return torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device).scatter_reduce_(
            dim, index_b, src, reduce="amax", include_self=True)
    raise ValueError(reduce)
IS_SHIM = True
