"""Restated slices of torch-geometric 2.0.2 / torch-scatter 2.0.7 (TEST SCAFFOLDING).

The reference pins these wheels (README.md:28-30) but they are not installable here, and
their source is not under /root/reference, so the arithmetic the reference's call sites
trigger is restated from the published PyG 2.0.2 semantics:

* ``GINConv(nn, eps, train_eps)``: ``nn(scatter_sum(x[src], dst) + (1 + eps) * x)``;
  call sites graph_classification/graph_neural_networks/models/gconv.py:197,212.
* ``RGCNConv(in, out, R, aggr='add')`` without bases/blocks: relations looped in order,
  ``out += scatter_sum(x[src_r], dst_r) @ weight[r]``; ``+ x @ root + bias``;
  call sites models/rgconv.py:96,121.
* ``global_add_pool / global_mean_pool / global_max_pool``: scatter over ``batch``.
* torch-scatter's CPU ``scatter_sum`` is ``zeros.scatter_add_`` (== ``index_add_`` rows).
"""
import math
import torch
from torch import nn


def scatter_sum(src, index, dim_size):
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def _num_graphs(batch):
    return int(batch.max().item()) + 1 if batch.numel() > 0 else 0


def global_add_pool(x, batch, size=None):
    return scatter_sum(x, batch, _num_graphs(batch) if size is None else size)


def global_mean_pool(x, batch, size=None):
    n = _num_graphs(batch) if size is None else size
    s = scatter_sum(x, batch, n)
    cnt = torch.bincount(batch, minlength=n).clamp(min=1).to(x.dtype).view(-1, 1)
    return s / cnt


def global_max_pool(x, batch, size=None):
    n = _num_graphs(batch) if size is None else size
    out = torch.full((n, x.size(1)), float("-inf"), dtype=x.dtype, device=x.device)
    return out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax", include_self=True)


class GINConv(nn.Module):
    def __init__(self, nn_module, eps=0.0, train_eps=False, **kw):
        super().__init__()
        self.nn = nn_module
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))

    def forward(self, x, edge_index, size=None):
        out = scatter_sum(x[edge_index[0]], edge_index[1], x.size(0))
        out = out + (1 + self.eps) * x
        return self.nn(out)


def _glorot(t):
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-stdv, stdv)


class RGCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, num_relations, num_bases=None, num_blocks=None,
                 aggr="mean", root_weight=True, bias=True, **kw):
        super().__init__()
        assert num_bases is None and num_blocks is None
        self.in_channels, self.out_channels, self.num_relations, self.aggr = (
            in_channels, out_channels, num_relations, aggr)
        self.weight = nn.Parameter(torch.Tensor(num_relations, in_channels, out_channels))
        self.root = nn.Parameter(torch.Tensor(in_channels, out_channels)) if root_weight else None
        self.bias = nn.Parameter(torch.Tensor(out_channels)) if bias else None
        _glorot(self.weight)
        if self.root is not None:
            _glorot(self.root)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, edge_index, edge_type=None):
        n = x.size(0)
        out = torch.zeros(n, self.out_channels, dtype=x.dtype, device=x.device)
        edge_type = edge_type.long()
        for r in range(self.num_relations):
            sel = edge_type == r
            src, dst = edge_index[0][sel], edge_index[1][sel]
            h = scatter_sum(x[src], dst, n)
            if self.aggr == "mean":
                cnt = torch.bincount(dst, minlength=n).clamp(min=1).to(x.dtype).view(-1, 1)
                h = h / cnt
            out = out + h @ self.weight[r]
        if self.root is not None:
            out = out + x @ self.root
        if self.bias is not None:
            out = out + self.bias
        return out


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not part of the restated hot path")


GCNConv = SAGEConv = FastRGCNConv = _Unavailable


def install(sys_modules):
    """Register stub modules ``torch_geometric{,.nn,.data,.datasets}``."""
    import types

    root = types.ModuleType("torch_geometric")
    nnm = types.ModuleType("torch_geometric.nn")
    for name in ("GINConv", "RGCNConv", "FastRGCNConv", "GCNConv", "SAGEConv",
                 "global_add_pool", "global_mean_pool", "global_max_pool"):
        setattr(nnm, name, globals()[name])
    datam = types.ModuleType("torch_geometric.data")

    def _no_network(*a, **k):
        raise RuntimeError("no network in this container")

    datam.download_url = _no_network
    datam.extract_zip = _no_network
    datam.InMemoryDataset = object
    dsm = types.ModuleType("torch_geometric.datasets")
    dsm.TUDataset = object
    root.nn, root.data, root.datasets = nnm, datam, dsm
    sys_modules["torch_geometric"] = root
    sys_modules["torch_geometric.nn"] = nnm
    sys_modules["torch_geometric.data"] = datam
    sys_modules["torch_geometric.datasets"] = dsm
