"""Restatement of ``torch-scatter==2.1.0`` and ``torch-cluster==1.6.0`` ops on the hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Both wheels are pinned in the
reference's ``env.yaml`` and absent from ``/root/reference`` and from this image.

torch_scatter (csrc/scatter.cpp, torch_scatter/scatter.py):
    ``scatter(src, index, dim, dim_size, reduce)`` with reduce in sum/add/mean;
    ``mean`` = sum / count.clamp(min=1) (empty rows stay 0).
    Call sites: ``tpscore.py:190,525``, ``conformer_utils.py:442``.

torch_cluster (csrc/cuda/radius_cuda.cu, torch_cluster/radius.py), CUDA semantics:
    ``radius(x, y, r, batch_x, batch_y, max_num_neighbors)``: one thread per query
    ``y_i`` walks the ``x_j`` of the same batch example in ascending ``j``, keeps pairs
    with ``sum_d (x_jd - y_id)^2 < r*r`` (fp32, strict) and stops after
    ``max_num_neighbors`` hits; returns ``[2, E] = (row=i over y, col=j over x)`` grouped by i.
    ``radius_graph(x, r, batch, loop=False, max_num_neighbors=32, flow='source_to_target')``
    = ``radius(x, x, r, batch, batch, max_num_neighbors + 1)`` then ``[col, row]`` with
    self pairs removed (torch_cluster/radius.py:81-128).
    Call sites: ``tpscore.py:586,613,655-660,721-723,747-749``.
"""
from __future__ import annotations

from typing import Optional

import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    dim_size = int(dim_size)
    res = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    res.index_add_(0, index.long(), src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index.long(), torch.ones(index.shape[0], dtype=src.dtype, device=src.device))
        cnt = cnt.clamp(min=1)
        return res / cnt.reshape((-1,) + (1,) * (src.dim() - 1))
    raise NotImplementedError(reduce)


def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


scatter_add = scatter_sum


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "mean")


def radius(x, y, r, batch_x: Optional[torch.Tensor] = None, batch_y: Optional[torch.Tensor] = None,
           max_num_neighbors: int = 32, num_workers: int = 1):
    """Vectorised per batch example; distance arithmetic in the dtype of ``x`` in the
    CUDA kernel's order ``((dx*dx) + dy*dy) + dz*dz`` without FMA contraction."""
    x = x.reshape(x.shape[0], -1)
    y = y.reshape(y.shape[0], -1)
    if batch_x is None:
        batch_x = torch.zeros(x.shape[0], dtype=torch.long)
    if batch_y is None:
        batch_y = torch.zeros(y.shape[0], dtype=torch.long)
    r2 = torch.tensor(r, dtype=x.dtype) * torch.tensor(r, dtype=x.dtype)
    rows, cols = [], []
    if y.shape[0] == 0 or x.shape[0] == 0:
        return torch.zeros(2, 0, dtype=torch.long)
    nb = int(max(batch_x.max(), batch_y.max())) + 1
    for b in range(nb):
        ix = torch.nonzero(batch_x == b).flatten()
        iy = torch.nonzero(batch_y == b).flatten()
        if ix.numel() == 0 or iy.numel() == 0:
            continue
        d = x[ix][None, :, :] - y[iy][:, None, :]  # [ny, nx, D]
        d2 = d[..., 0] * d[..., 0]
        for k in range(1, d.shape[-1]):
            d2 = d2 + d[..., k] * d[..., k]
        hit = d2 < r2
        rank = torch.cumsum(hit.to(torch.long), dim=1)
        hit = hit & (rank <= max_num_neighbors)
        yi, xj = torch.nonzero(hit, as_tuple=True)  # row-major: grouped by y, ascending x
        rows.append(iy[yi])
        cols.append(ix[xj])
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long)
    row, col = torch.cat(rows), torch.cat(cols)
    # examples are contiguous in PyG batches, so concatenation order == ascending row
    return torch.stack([row, col], dim=0)


def radius_graph(x, r, batch=None, loop: bool = False, max_num_neighbors: int = 32,
                 flow: str = "source_to_target", num_workers: int = 1):
    assert flow in ("source_to_target", "target_to_source")
    edge_index = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    if flow == "source_to_target":
        row, col = edge_index[1], edge_index[0]
    else:
        row, col = edge_index[0], edge_index[1]
    if not loop:
        mask = row != col
        row, col = row[mask], col[mask]
    return torch.stack([row, col], dim=0)
