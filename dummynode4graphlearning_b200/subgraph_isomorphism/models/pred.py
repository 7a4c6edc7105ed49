"""Counting / matching head (``subgraph_isomorphism/models/pred.py:17-237``).

Same modules, parameters (``p_fc, g_fc, pred_fc1, pred_fc2[, weight_fc1, weight_fc2]``) and padded
semantics as the reference: the head sees LEFT-PADDED (B, L, rep) tensors whose padded and masked rows
are zero, projects every row (so padded rows contribute the projection's bias) and pools over the whole
padded axis.  The padded tensors are produced by one CUDA kernel
(``ops.pad_segments``) instead of the reference's per-graph Python loop; the dense projections are
library GEMMs.
"""
import torch as th
import torch.nn as nn

from ... import ops
from ..utils import init_module, map_activation_str_to_layer


class PredictNet(nn.Module):
    def __init__(self, input_dim, hidden_dim, act_func="relu", dropout=0.0, return_weights=False):
        super().__init__()
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.act = map_activation_str_to_layer(act_func)
        self.drop = nn.Dropout(dropout)
        self.p_fc = ops.Linear(input_dim, hidden_dim)
        self.g_fc = ops.Linear(input_dim, hidden_dim)
        self.pred_fc1 = nn.Linear(hidden_dim * 4 + 4, hidden_dim)
        self.pred_fc2 = nn.Linear(hidden_dim + 4, 1)
        if return_weights:
            self.weight_fc1 = nn.Linear(hidden_dim * 4 + 2, hidden_dim)
            self.weight_fc2 = nn.Linear(hidden_dim + 2, 1)
        else:
            self.weight_fc1 = self.weight_fc2 = None
        for m, how in ((self.p_fc, "normal"), (self.g_fc, "normal"), (self.pred_fc1, "normal"),
                       (self.pred_fc2, "zero")):           # pred.py:47-50 (fc2 zero-init: outputs 0 at init)
            init_module(m, activation=act_func, init=how)
        if return_weights:
            init_module(self.weight_fc1, activation=act_func, init="normal")
            init_module(self.weight_fc2, activation=act_func, init="zero")

    def agg_graph(self, g_rep, g_mask=None):
        raise NotImplementedError

    def forward(self, p_rep, p_mask, g_rep, g_mask):
        bsz, g_len = p_mask.size(0), g_mask.size(1)
        pl = p_mask.float().sum(dim=1).view(bsz, 1)
        gl = g_mask.float().sum(dim=1).view(bsz, 1)
        pl_inv, gl_inv = 1.0 / pl, 1.0 / gl
        if p_rep.dim() == 2:
            p_vec = p_rep
        elif p_rep.dim() == 3:
            p_vec = self.agg_graph(self.drop(self.p_fc(p_rep)), p_mask)
        else:
            raise ValueError
        g = self.drop(self.g_fc(g_rep))
        if self.weight_fc1 is not None:   # per-node matching weights (pred.py:114-136)
            p = p_vec.unsqueeze(1).expand(bsz, g_len, -1)
            plx = pl.expand(bsz, g_len).unsqueeze(-1)
            plix = pl_inv.expand(bsz, g_len).unsqueeze(-1)
            w = self.act(self.weight_fc1(th.cat([p, g, g - p, g * p, plx, plix], dim=2)))
            w = self.weight_fc2(th.cat([w, plx, plix], dim=2)).squeeze(-1)
        else:
            w = None
        gv = self.agg_graph(g)
        y = th.cat([p_vec, gv, gv - p_vec, gv * p_vec, pl, gl, pl_inv, gl_inv], dim=1)
        y = self.act(self.pred_fc1(y))
        y = self.pred_fc2(th.cat([y, pl, gl, pl_inv, gl_inv], dim=1))
        return y, w


class MeanPredictNet(PredictNet):
    def agg_graph(self, g_rep, g_mask=None):
        return th.mean(g_rep, dim=1)


class SumPredictNet(PredictNet):
    def agg_graph(self, g_rep, g_mask=None):
        return th.sum(g_rep, dim=1)


class MaxPredictNet(PredictNet):
    def agg_graph(self, g_rep, g_mask=None):
        return th.max(g_rep, dim=1)[0]
