"""Relational GCN on the B200 aggregation kernels.

Drop-in for ``subgraph_isomorphism/models/rgcn.py``: ``RGCNLayer`` (:16-213) and ``RGCN`` (:213-300) keep the
constructor arguments, ``forward(g, node_feat, edge_type) -> (node_out, edge_type)`` and the parameter names
``weight, w_comp, loop_weight, bias, bn.*`` (module names ``graph_rgcn_(i)``).

Same restructuring as ``rgin.py`` here: the per-relation projections are taken once per NODE
(``table[u, r, :] = h[u] @ W_r``) and the message sum is the CSR gather-sum kernel (K1) over rows
``src(e) * R + type(e)`` of that table.  The edge normalisation of the reference (rgcn.py:134-165) factorises over the
endpoints, so it never touches an edge:

  edge_norm "in":    msg_e * innorm[dst(e)]                         -> scale row v of the aggregate by innorm[v]
  edge_norm "both":  msg_e * sqrt(outnorm[src(e)] * innorm[dst(e)])  -> scale h[u] by sqrt(outnorm[u]) before the table
                                                                       GEMM and row v of the aggregate by sqrt(innorm[v])
with innorm = 1 / (in_deg + 1), outnorm = 1 / (out_deg + 1) (self loops are always on, rgcn.py:222-239).
"""
import torch as th
import torch.nn as nn

from ... import ops
from ..utils import init_weight, map_activation_str_to_layer
from .basemodel import GraphAdjModel
from .rgin import RGINLayer, relation_csr


class RGCNLayer(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_rels=1, regularizer="basis", num_bases=-1, edge_norm="in",
                 self_loop=True, bias=True, batch_norm=False, act_func="relu", dropout=0.0):
        super().__init__()
        assert regularizer in ["none", "basis", "bdd"]
        assert edge_norm in ["none", "in", "both"]
        self.input_dim, self.hidden_dim, self.num_rels, self.regularizer = input_dim, hidden_dim, num_rels, regularizer
        if regularizer == "none" or num_bases is None or num_bases > num_rels or num_bases <= 0:
            self.num_bases = num_rels
        else:
            self.num_bases = num_bases
        self.edge_norm = edge_norm
        if self_loop:
            self.loop_weight = nn.Parameter(th.Tensor(input_dim, hidden_dim))
        else:
            self.register_parameter("loop_weight", None)
        if bias:
            self.bias = nn.Parameter(th.Tensor(hidden_dim))
        else:
            self.register_parameter("bias", None)
        self.bn = nn.BatchNorm1d(hidden_dim) if batch_norm else None
        self.act = map_activation_str_to_layer(act_func)
        self.drop = nn.Dropout(dropout)
        if regularizer in ("none", "basis"):
            self.weight = nn.Parameter(th.Tensor(self.num_bases, input_dim, hidden_dim))
            if self.num_bases < self.num_rels:
                self.w_comp = nn.Parameter(th.Tensor(self.num_rels, self.num_bases))
            else:
                self.register_parameter("w_comp", None)
        else:
            if input_dim % self.num_bases != 0 or hidden_dim % self.num_bases != 0:
                raise ValueError("Feature size must be a multiplier of num_bases (%d)." % self.num_bases)
            self.weight = nn.Parameter(
                th.Tensor(self.num_rels, self.num_bases * (input_dim // self.num_bases) * (hidden_dim // self.num_bases)))
            self.register_parameter("w_comp", None)
        init_weight(self.weight, activation=act_func, init="uniform")
        if self.w_comp is not None:
            init_weight(self.w_comp, activation=act_func, init="uniform")
        if self_loop:
            init_weight(self.loop_weight, activation=act_func, init="uniform")
        if bias:
            nn.init.zeros_(self.bias)

    self_loop = RGINLayer.self_loop
    relation_weights = RGINLayer.relation_weights

    def _norms(self, g):
        """(innorm, outnorm) as (N, 1) float columns, cached on the graph (rgcn.py:134-154)."""
        def make():
            one = 1.0 if self.self_loop else 0.0
            ind, outd = g.in_degrees().float(), g.out_degrees().float()
            if self.self_loop:
                return (1.0 / (ind + one)).view(-1, 1), (1.0 / (outd + one)).view(-1, 1)
            return ((1.0 / ind).masked_fill(ind == 0, 0.0).view(-1, 1), (1.0 / outd).masked_fill(outd == 0, 0.0).view(-1, 1))
        return g.cached(("rgcn_norms", self.self_loop), make)

    def forward(self, g, node_feat, edge_type):
        fwd, bwd = relation_csr(g, edge_type, self.num_rels)
        innorm = outnorm = None
        if self.edge_norm != "none":
            innorm, outnorm = self._norms(g)
        src_feat = node_feat * outnorm.sqrt() if self.edge_norm == "both" else node_feat
        table = ops.matmul_xw(src_feat, self.relation_weights()).view(-1, self.hidden_dim)    # (N*R, H)
        out = ops.spmm_sum(table, fwd, bwd)                                                 # fn.sum, rgcn.py:97
        if self.edge_norm == "in":
            out = out * innorm
        elif self.edge_norm == "both":
            out = out * innorm.sqrt()
        if self.self_loop:                                                                    # rgcn.py:170-180
            loop = ops.matmul_xw(node_feat, self.loop_weight)
            if self.edge_norm == "in":
                loop = loop * innorm
            elif self.edge_norm == "both":
                loop = loop * (innorm * outnorm).sqrt()
            out = out + loop
        if self.bias is not None:
            out = out + self.bias
        if self.bn is not None:
            out = self.bn(out)
        out = self.drop(self.act(out))
        return out, edge_type

    def get_output_dim(self):
        return self.hidden_dim

    def extra_repr(self):
        return "in=%d, out=%d, num_rels=%d, regularizer=%s, num_bases=%d, edge_norm=%s, self_loop=%s, bias=%s" % (
            self.input_dim, self.hidden_dim, self.num_rels, self.regularizer, self.num_bases, self.edge_norm,
            self.self_loop, self.bias is not None)


class RGCN(GraphAdjModel):
    def create_rep_net(self, type, **kw):
        if type == "graph":
            num_layers, num_rels = kw.get("rep_num_graph_layers", 1), self.max_ngel
        else:
            if self.share_rep_net:
                return self.g_rep_net
            num_layers, num_rels = kw.get("rep_num_pattern_layers", 1), self.max_npel
        layers = nn.ModuleList()
        for i in range(num_layers):
            layers.add_module(
                "%s_rgcn_(%d)" % (type, i),
                RGCNLayer(self.hid_dim, self.hid_dim, num_rels=num_rels,
                          regularizer=kw.get("rep_rgcn_regularizer", "basis"), num_bases=kw.get("rep_rgcn_num_bases", -1),
                          edge_norm=kw.get("rep_rgcn_edge_norm", "in"), batch_norm=kw.get("rep_rgcn_batch_norm", False),
                          act_func=kw.get("rep_act_func", "relu"), dropout=kw.get("rep_dropout", 0.0)))
        return nn.ModuleDict({"rgcn": layers})

    def _run(self, net, g, h, gate):
        etype = g.edata["label"]
        for layer in net["rgcn"]:
            o, etype = layer(g, h, etype)
            if gate is not None:
                o = o * gate
            h = h + o if (self.rep_residual and h.size() == o.size()) else o
        return h

    def get_pattern_rep(self, pattern, p_emb, mask=None):
        if mask is not None:   # rgcn.py:253-259: the masked variant has no residual
            zero = ~mask
            h = p_emb.masked_fill(zero, 0.0)
            etype = pattern.edata["label"]
            for layer in self.p_rep_net["rgcn"]:
                o, etype = layer(pattern, h, etype)
                h = o.masked_fill(zero, 0.0)
            return h
        return self._run(self.p_rep_net, pattern, p_emb, None)

    def get_graph_rep(self, graph, g_emb, mask=None, gate=None):
        if mask is None and gate is None:
            return self._run(self.g_rep_net, graph, g_emb, None)
        if gate is None:
            gate = mask.float()
        elif mask is not None:
            gate = mask.float() * gate
        return self._run(self.g_rep_net, graph, g_emb * gate, gate)
