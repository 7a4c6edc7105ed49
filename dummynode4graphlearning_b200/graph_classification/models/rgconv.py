"""Relational GIN classifier (``graph_classification/graph_neural_networks/models/rgconv.py::RGIN`` :53-126) and
the PyG 2.0.2 ``RGCNConv(in, out, R, aggr='add')`` it calls (:96,121).

The reference loops over relations in Python, masking the edge list and running one propagate + GEMM per
relation.  Here: one dense GEMM builds the (N, R*D) per-relation table, one K1 launch sums row
``src*R + type`` over every in-edge.  ``args.nhid`` is read like the reference does (:64); when absent it falls
back to ``args.hidden_dim`` (the reference's main.py never defines nhid -- SURVEY.md App. A-15).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ..data import structure_of
from .gconv import _gin_mlp, global_add_pool, global_mean_pool


class RGCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, num_relations, aggr="add", root_weight=True, bias=True, **kw):
        super().__init__()
        if aggr != "add":
            raise NotImplementedError("only aggr='add' (the RGIN call site) is on the hot path")
        self.in_channels, self.out_channels, self.num_relations = in_channels, out_channels, num_relations
        self.weight = nn.Parameter(torch.Tensor(num_relations, in_channels, out_channels))
        if root_weight:
            self.root = nn.Parameter(torch.Tensor(in_channels, out_channels))
        else:
            self.register_parameter("root", None)
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        for w in (self.weight, self.root):   # glorot
            if w is not None:
                a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
                w.data.uniform_(-a, a)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, edge_index, edge_type, structure=None):
        if structure is None:
            from ..data import GraphStructure
            structure = GraphStructure(edge_index, x.size(0))
        R, I, O = self.num_relations, self.in_channels, self.out_channels
        fwd, bwd = structure.relation_csr(edge_type, R)
        table = ops.matmul_xw(x, self.weight.permute(1, 0, 2).reshape(I, R * O)).view(-1, O)
        out = ops.spmm_sum(table, fwd, bwd)
        if self.root is not None:
            out = out + ops.matmul_xw(x, self.root)
        if self.bias is not None:
            out = out + self.bias
        return out


class RGIN(torch.nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.num_features = args.num_features
        self.nhid = getattr(args, "nhid", None) or args.hidden_dim
        self.num_classes, self.dropout, self.num_relations = args.num_classes, args.dropout_ratio, args.num_relations
        config = args.additional if args.additional else {"num_layers": 2}
        agg = config.get("aggregation", "sum")
        if agg == "sum":
            self.pooling = global_add_pool
        elif agg == "mean":
            self.pooling = global_mean_pool
        self.embeddings_dim = [self.nhid for _ in range(config.get("num_layers", 2))]
        self.no_layers = len(self.embeddings_dim)
        nns, convs, linears = [], [], []
        for layer, out_dim in enumerate(self.embeddings_dim):
            if layer == 0:
                self.first_h = _gin_mlp(self.num_features, out_dim)
            else:
                nns.append(_gin_mlp(self.embeddings_dim[layer - 1], out_dim))
                convs.append(RGCNConv(self.nhid, self.nhid, self.num_relations, aggr="add"))
            linears.append(ops.Linear(out_dim, self.num_classes))
        if ("weight_reg" in config) and (config["weight_reg"] > 1.1):
            with torch.no_grad():
                for conv in convs:
                    conv.weight.div_(config["weight_reg"])
        self.nns = nn.ModuleList(nns)
        self.convs = nn.ModuleList(convs)
        self.linears = nn.ModuleList(linears)

    def forward(self, data):
        x, edge_attr = data.x, data.edge_attr
        s = structure_of(data)
        if edge_attr is not None:
            edge_type = edge_attr.max(dim=1)[1]                      # rgconv.py:110-111
        else:
            edge_type = torch.zeros(data.edge_index.size(1), dtype=torch.long, device=x.device)
        out = 0
        for layer in range(self.no_layers):
            if layer == 0:
                x = ops.apply_gin_mlp(self.first_h, x)
                out += F.dropout(self.pooling(self.linears[layer](x), data.batch, node_ptr=s.node_ptr), p=self.dropout)
            else:
                x = self.convs[layer - 1](x, data.edge_index, edge_type, structure=s)
                x = ops.apply_gin_mlp(self.nns[layer - 1], x)
                out += F.dropout(self.linears[layer](self.pooling(x, data.batch, node_ptr=s.node_ptr)),
                                 p=self.dropout, training=self.training)
        return F.log_softmax(out, dim=-1)
