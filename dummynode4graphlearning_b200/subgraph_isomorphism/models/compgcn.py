"""CompGCN on the B200 kernels.

Drop-in for ``subgraph_isomorphism/models/compgcn.py``: ``CompGCNLayer`` (:101-287) and ``CompGCN`` (:290-385) keep the
constructor arguments, ``forward(graph, node_feat, edge_feat) -> (node_out, edge_out)`` and the parameter names
``loop_weight, bias, bn.*, in_weight, out_weight, rel_weight, loop_rel`` (module names ``graph_compgcn_(i)``).

Restructured like DMPNN here (ops.dmp_node_agg): the reference multiplies every gathered per-edge composition by
``in_weight`` / ``out_weight`` (two (E, D) x (D, D) matmuls + two masked_fill copies, compgcn.py:228-236).  By linearity

    agg[v] = (sum_{e in in(v), !rev} n_e c_e) @ W_in + (sum_{e in in(v), rev} n_e c_e) @ W_out,   c_e = comp(h[src e], ef[e])

and the edge norm factorises over the endpoints, n_e = a[src e] * b[dst e] (compgcn.py:190-209):
``in``: (1, innorm), ``out``: (outnorm, 1), ``both``: (sqrt(outnorm), sqrt(innorm)).  So the per-edge work is one
streaming composition kernel (``dn4gl_comp_edge_f32``: sub / mult), the reduce is the K4 segment sum and the weights are
ONE node-level GEMM on ``[S_rev | S_fwd]``.  ``corr`` (circular correlation through rfft/irfft, compgcn.py:219-223) keeps
torch's FFT for the composition and uses the same segment sum.
"""
import torch as th
import torch.nn as nn

from ... import ops
from ..utils import init_weight, map_activation_str_to_layer
from .dmpnn import DMPNN


def _circular_correlation(head, relation):
    """irfft(conj(rfft(head)) * rfft(relation)) along the feature axis (compgcn.py:219-223)."""
    n = head.size(-1)
    return th.fft.irfft(th.conj(th.fft.rfft(head, dim=-1)) * th.fft.rfft(relation, dim=-1), n=n, dim=-1)


class CompGCNLayer(nn.Module):
    def __init__(self, input_dim, hidden_dim, self_loop=True, comp_opt="mult", edge_norm="both", bias=True,
                 batch_norm=False, act_func="relu", dropout=0.0):
        super().__init__()
        assert edge_norm in ["none", "in", "out", "both"]
        self.input_dim, self.hidden_dim, self.edge_norm, self.comp_opt = input_dim, hidden_dim, edge_norm, comp_opt
        self.num_rels = 3 if self_loop else 2
        if self_loop:
            self.loop_weight = nn.Parameter(th.Tensor(input_dim, hidden_dim))
        else:
            self.register_parameter("loop_weight", None)
        if bias:
            self.bias = nn.Parameter(th.Tensor(hidden_dim))
        else:
            self.register_parameter("bias", None)
        self.bn = nn.BatchNorm1d(hidden_dim) if batch_norm else None
        self.in_weight = nn.Parameter(th.Tensor(input_dim, hidden_dim))
        self.out_weight = nn.Parameter(th.Tensor(input_dim, hidden_dim))
        self.rel_weight = nn.Parameter(th.Tensor(input_dim, hidden_dim))
        if self_loop:
            self.loop_rel = nn.Parameter(th.Tensor(1, input_dim))
        else:
            self.register_parameter("loop_rel", None)
        self.act = map_activation_str_to_layer(act_func)
        self.drop = nn.Dropout(dropout)
        for w in (self.in_weight, self.out_weight, self.rel_weight):       # same order as compgcn.py:152-160 (RNG parity)
            init_weight(w, activation=act_func, init="uniform")
        if self_loop:
            init_weight(self.loop_weight, activation=act_func, init="uniform")
            init_weight(self.loop_rel, activation=act_func, init="uniform")
        if bias:
            nn.init.zeros_(self.bias)

    @property
    def self_loop(self):
        return getattr(self, "loop_weight", None) is not None

    def _norms(self, g):
        """(innorm, outnorm) as (N,) float vectors, cached on the graph (compgcn.py:177-198)."""
        def make():
            ind, outd = g.in_degrees().float(), g.out_degrees().float()
            if self.self_loop:
                return 1.0 / (ind + 1.0), 1.0 / (outd + 1.0)
            return (1.0 / ind).masked_fill(ind == 0, 1.0), (1.0 / outd).masked_fill(outd == 0, 1.0)
        return g.cached(("compgcn_norms", self.self_loop), make)

    def _comp(self, head, relation):
        if self.comp_opt == "sub":
            return head - relation
        if self.comp_opt == "mult":
            return head * relation
        if self.comp_opt == "corr":
            return _circular_correlation(head, relation.expand_as(head))
        raise NotImplementedError(self.comp_opt)

    def forward(self, graph, node_feat, edge_feat):
        a = b = None                                   # n_e = a[src e] * b[dst e]
        if self.edge_norm != "none":
            innorm, outnorm = self._norms(graph)
            if self.edge_norm == "in":
                b = innorm
            elif self.edge_norm == "out":
                a = outnorm
            else:
                a, b = outnorm.sqrt(), innorm.sqrt()
        # ---- node stream: update_all(message :225-240, fn.sum :163, update :242-261)
        if self.comp_opt in ("sub", "mult"):
            C = ops.comp_edge(node_feat, edge_feat, graph, ops.COMP_SUB if self.comp_opt == "sub" else ops.COMP_MULT, a)
        else:
            src = graph.src.long()
            C = self._comp(node_feat.index_select(0, src), edge_feat)
            if a is not None:
                C = C * a.index_select(0, src).view(-1, 1)
        S = ops.dmp_node_agg(C, graph)                                       # (N, 2D) = [S_rev | S_fwd]
        if b is not None:
            S = S * b.view(-1, 1)
        out = ops.matmul_xw(S, th.cat([self.out_weight, self.in_weight], dim=0))
        if self.self_loop:
            loop = ops.matmul_xw(self._comp(node_feat, self.loop_rel), self.loop_weight)
            out = (out + loop) * 0.3333333
        else:
            out = out * 0.5
        if self.bias is not None:
            out = out + self.bias
        if self.bn is not None:
            out = self.bn(out)
        out = self.drop(self.act(out))
        # ---- edge stream: apply_edges(:263-266)
        return out, ops.matmul_xw(edge_feat, self.rel_weight)

    def extra_repr(self):
        return "in=%s, out=%s, comp_opt=%s, edge_norm=%s, self_loop=%s, bias=%s" % (
            self.input_dim, self.hidden_dim, self.comp_opt, self.edge_norm, self.self_loop, self.bias is not None)

    def get_output_dim(self):
        return self.hidden_dim


class CompGCN(DMPNN):
    """same layer loop / masking / gating / residual wiring as DMPNN (compgcn.py:323-385 == dmpnn.py:215-277)."""
    rep_key = "compgcn"

    def create_rep_net(self, type, **kw):
        if type == "graph":
            num_layers = kw.get("rep_num_graph_layers", 1)
        else:
            if self.share_rep_net:
                return self.g_rep_net
            num_layers = kw.get("rep_num_pattern_layers", 1)
        layers = nn.ModuleList()
        for i in range(num_layers):
            layers.add_module(
                "%s_compgcn_(%d)" % (type, i),
                CompGCNLayer(self.hid_dim, self.hid_dim, comp_opt=kw.get("rep_compgcn_comp_opt", "mult"),
                             edge_norm=kw.get("rep_compgcn_edge_norm", "none"),
                             batch_norm=kw.get("rep_compgcn_batch_norm", False),
                             act_func=kw.get("rep_act_func", "relu"), dropout=kw.get("rep_dropout", 0.0)))
        return nn.ModuleDict({"compgcn": layers})
