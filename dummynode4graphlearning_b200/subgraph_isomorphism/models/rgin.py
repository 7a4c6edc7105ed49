"""Relational GIN on the B200 aggregation kernels.

Drop-in for ``subgraph_isomorphism/models/rgin.py``: ``RGINLayer`` (:16-172) and ``RGIN`` (:175-260)
keep the constructor arguments, ``forward(g, node_feat, edge_type) -> (node_out, edge_type)`` and the
parameter names ``weight, w_comp, loop_weight, bias, mlp.*`` (module names ``graph_rgin_(i)``).

Restructuring (verified there to 3e-7): the reference materialises a weight
matrix PER EDGE (``weight.index_select(0, etype)`` + ``bmm``, :102-120, traffic ~ E*D^2).  Here the
per-relation projections are taken once per NODE by one dense GEMM,
``table[u, r, :] = h[u] @ W_r`` (N x R*D), and the message sum becomes the generic CSR gather-sum
kernel (K1) reading row ``src(e)*R + type(e)`` of that table for every in-edge e -- traffic ~ E*D.
The backward is the same kernel on the transposed (N*R rows) CSR followed by the GEMM adjoints.
"""
import torch as th
import torch.nn as nn

from ... import ops
from ...graph import build_csr, CSR
from ..utils import init_weight, map_activation_str_to_layer
from .basemodel import GraphAdjModel


def relation_csr(g, edge_type, num_rels):
    """forward / transposed CSR pair whose columns address the (N*R, D) per-relation table."""

    def make():
        base = g.csr_in
        et = edge_type.to(th.int32)
        col = (base.col.long() * num_rels + et.long()[base.eid.long()]).to(th.int32)
        fwd = CSR(base.row_ptr, col, base.eid, base.n_rows, base.nnz)
        fwd.heavy_rows, fwd.heavy_count, fwd.heavy_thr = base.heavy_rows, base.heavy_count, base.heavy_thr
        tkey = (g.src.long() * num_rels + et.long()).to(th.int32)
        bwd = build_csr(tkey, g.dst, g.number_of_nodes() * num_rels)
        return fwd, bwd

    from ...graph import cached_for_tensor
    return cached_for_tensor(g._cache, ("rel_csr", int(num_rels)), edge_type, int(num_rels), make)


class RGINLayer(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_rels=1, regularizer="basis", num_bases=-1, num_mlp_layers=2,
                 self_loop=True, bias=True, batch_norm=False, act_func="relu", dropout=0.0):
        super().__init__()
        assert regularizer in ["none", "basis", "bdd"]
        self.input_dim, self.hidden_dim, self.num_rels, self.regularizer = input_dim, hidden_dim, num_rels, regularizer
        if regularizer == "none" or num_bases is None or num_bases > num_rels or num_bases <= 0:
            self.num_bases = num_rels
        else:
            self.num_bases = num_bases
        if self_loop:
            self.loop_weight = nn.Parameter(th.Tensor(input_dim, hidden_dim))
        else:
            self.register_parameter("loop_weight", None)
        if bias:
            self.bias = nn.Parameter(th.Tensor(hidden_dim))
        else:
            self.register_parameter("bias", None)
        mlp = []
        for i in range(num_mlp_layers):
            mlp.append(ops.Linear(hidden_dim, hidden_dim))
            if i != num_mlp_layers - 1:
                if batch_norm:
                    mlp.append(nn.BatchNorm1d(hidden_dim))
                mlp.append(map_activation_str_to_layer(act_func))
        self.mlp = nn.Sequential(*mlp)
        self.act = map_activation_str_to_layer(act_func)
        self.drop = nn.Dropout(dropout)

        if regularizer in ("none", "basis"):
            self.weight = nn.Parameter(th.Tensor(self.num_bases, input_dim, hidden_dim))
            if self.num_bases < self.num_rels:
                self.w_comp = nn.Parameter(th.Tensor(self.num_rels, self.num_bases))
            else:
                self.register_parameter("w_comp", None)
        else:  # bdd: block-diagonal, num_bases blocks of (in/nb) x (out/nb) per relation
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

    @property
    def self_loop(self):
        return self.loop_weight is not None

    def relation_weights(self):
        """(input_dim, num_rels * hidden_dim): column block r holds W_r."""
        R, I, H = self.num_rels, self.input_dim, self.hidden_dim
        if self.regularizer in ("none", "basis"):
            w = self.weight
            if self.num_bases < R:   # W_r = sum_b w_comp[r, b] V_b          (rgin.py:103-106)
                w = th.matmul(self.w_comp, w.view(self.num_bases, I * H)).view(R, I, H)
        else:                        # block-diagonal W_r                     (rgin.py:114-118)
            nb = self.num_bases
            si, so = I // nb, H // nb
            blocks = self.weight.view(R, nb, si, so)
            eye = th.eye(nb, dtype=blocks.dtype, device=blocks.device)
            w = th.einsum("rbio,bc->rbico", blocks, eye).reshape(R, I, H)
        return w.permute(1, 0, 2).reshape(I, R * H)

    def forward(self, g, node_feat, edge_type):
        fwd, bwd = relation_csr(g, edge_type, self.num_rels)
        table = ops.matmul_xw(node_feat, self.relation_weights()).view(-1, self.hidden_dim)   # (N*R, H)
        out = ops.spmm_sum(table, fwd, bwd)                                                 # fn.sum, rgin.py:98
        if self.self_loop:
            out = out + ops.matmul_xw(node_feat, self.loop_weight)
        if self.bias is not None:
            out = out + self.bias
        if len(self.mlp) == 0:
            out = self.act(out)
        elif out.is_cuda and ops.mlp2_fusable(self.mlp):
            out = ops.mlp2(self.mlp, out)          # Linear, act, Linear on the tensor cores (csrc/mlp_tc.cu)
        else:
            out = self.mlp(out)
        out = self.act(out)   # the reference applies the activation once more after the MLP (rgin.py:147-151)
        out = self.drop(out)
        return out, edge_type

    def get_output_dim(self):
        return self.hidden_dim

    def extra_repr(self):
        # rgin.py:169 references a non-existent self.edge_norm (print(model) raises there); not replicated.
        return "in=%d, out=%d, num_rels=%d, regularizer=%s, num_bases=%d, self_loop=%s, bias=%s" % (
            self.input_dim, self.hidden_dim, self.num_rels, self.regularizer, self.num_bases, self.self_loop,
            self.bias is not None)


class RGIN(GraphAdjModel):
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
                "%s_rgin_(%d)" % (type, i),
                RGINLayer(self.hid_dim, self.hid_dim, num_rels=num_rels,
                          regularizer=kw.get("rep_rgin_regularizer", "basis"),
                          num_bases=kw.get("rep_rgin_num_bases", -1),
                          num_mlp_layers=kw.get("rep_rgin_num_mlp_layers", 2),
                          batch_norm=kw.get("rep_rgin_batch_norm", False),
                          act_func=kw.get("rep_act_func", "relu"), dropout=kw.get("rep_dropout", 0.0)))
        return nn.ModuleDict({"rgin": layers})

    def _run(self, net, g, h, gate):
        etype = g.edata["label"]
        for layer in net["rgin"]:
            o, etype = layer(g, h, etype)
            if gate is not None:
                o = o * gate
            h = h + o if (self.rep_residual and h.size() == o.size()) else o
        return h

    def get_pattern_rep(self, pattern, p_emb, mask=None):
        if mask is not None:  # rgin.py:215-221: masked variant has no residual
            zero = ~mask
            h = p_emb.masked_fill(zero, 0.0)
            etype = pattern.edata["label"]
            for layer in self.p_rep_net["rgin"]:
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
