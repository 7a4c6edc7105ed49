"""Counting / matching head (``subgraph_isomorphism/models/pred.py:17-237``).

Same modules, parameters (``p_fc, g_fc, pred_fc1, pred_fc2[, weight_fc1, weight_fc2]``) and padded
semantics as the reference: the head sees LEFT-PADDED (B, L, rep) tensors whose padded and masked rows
are zero, projects every row (so padded rows contribute the projection's bias) and pools over the whole
padded axis.

``forward`` takes those padded tensors (``ops.pad_segments`` builds them in one kernel).  The models call
``forward_ragged`` instead whenever the pooling is a sum or a mean and dropout is inactive: every padded or masked row of
the reference's tensor is a zero row whose projection is exactly the bias, so  sum_rows (W x_r + b)  over the padded axis
=  sum over the graph's unmasked rows of W x_r  +  L b  -- the projection of the N real rows and a masked segment sum (K3)
replace the (B, L, rep) tensor and the projection of every padded row (pred.py:111,142,215-216; ``L`` is the batch-wide
padded length, so the batch-composition dependence of App. A-7 is kept).
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
        self.pred_fc1 = ops.Linear(hidden_dim * 4 + 4, hidden_dim)
        self.pred_fc2 = ops.Linear(hidden_dim + 4, 1)
        if return_weights:
            self.weight_fc1 = ops.Linear(hidden_dim * 4 + 2, hidden_dim)
            self.weight_fc2 = ops.Linear(hidden_dim + 2, 1)
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


    # ---- unpadded evaluation ------------------------------------------------------------------------------------
    ragged_pool = None      # "sum" | "mean" for heads whose pooling commutes with the projection

    def supports_ragged(self):
        return self.ragged_pool is not None and not (self.training and self.drop.p > 0)

    def _pooled(self, fc, x, seg_ptr, drop, L):
        """agg_graph(fc(padded rows)) without the padded tensor.  The projection stays BEFORE the sum, as in the reference:
        summing the raw rows first is algebraically the same but puts degree / multi-hot columns of magnitude ~1e5 through
        the projection's cancellations (measured: 2e-5 on the DMPNN 'large' loss instead of < 1e-6).  So: project the
        N real rows (not B * L padded ones), masked segment sum of the projections, + L biases (every padded or masked
        row of the reference's tensor is a zero row and contributes exactly the bias)."""
        h = ops.linear(x, fc.weight, None)                    # (N, H)
        v = ops.segment_sum(h, seg_ptr, drop) + float(L) * fc.bias
        return v / float(L) if self.ragged_pool == "mean" else v

    def forward_ragged(self, p_x, p_ptr, p_drop, p_mask, Lp, g_x, g_ptr, g_drop, g_mask, Lg):
        """p_x (Np, rep) / g_x (Ng, rep): flat readout rows; *_ptr int32 (B + 1) row offsets; *_drop (N,) bool rows the
        padded path zeroes (dummy nodes, reversed edges) or None; *_mask (B, L) bool as handed to ``forward``; Lp / Lg the
        padded lengths.  Same values as ``forward`` on the padded tensors (up to fp32 summation order); ``w`` is defined on
        the unpadded positions (the padded ones are don't-care: the loss masks them, train.py:783-784)."""
        bsz = p_mask.size(0)
        pl = p_mask.float().sum(dim=1).view(bsz, 1)
        gl = g_mask.float().sum(dim=1).view(bsz, 1)
        pl_inv, gl_inv = 1.0 / pl, 1.0 / gl
        p_vec = self._pooled(self.p_fc, p_x, p_ptr, p_drop, Lp)
        gv = self._pooled(self.g_fc, g_x, g_ptr, g_drop, Lg)
        w = None
        if self.weight_fc1 is not None:
            gx = g_x if g_drop is None else g_x.masked_fill(g_drop.view(-1, 1), 0.0)
            g = ops.linear(gx, self.g_fc.weight, self.g_fc.bias)                      # (Ng, H), per node
            seg = ops.segment_ids(g_ptr, g_x.size(0)).long()
            p, plx, plix = p_vec[seg], pl[seg], pl_inv[seg]
            wr = self.act(self.weight_fc1(th.cat([p, g, g - p, g * p, plx, plix], dim=1)))
            wr = self.weight_fc2(th.cat([wr, plx, plix], dim=1))                      # (Ng, 1)
            # padded positions: the reference evaluates the same formula on zero rows (g = bias), one value per graph.
            # The loss zeroes those positions IN PLACE outside autograd (train.py:783-784) but keeps their gradient path
            # (the match regulariser relu(pred_v - pred_c) sees 0 - pred_c there), so they must exist with their producer.
            gb = self.g_fc.bias.view(1, -1).expand(bsz, -1)
            wp = self.act(self.weight_fc1(th.cat([p_vec, gb, gb - p_vec, gb * p_vec, pl, pl_inv], dim=1)))
            wp = self.weight_fc2(th.cat([wp, pl, pl_inv], dim=1))                     # (B, 1)
            ones = th.ones((g_x.size(0), 1), dtype=wr.dtype, device=wr.device)
            is_pad = 1.0 - ops.pad_segments(ones, g_ptr, int(Lg)).view(bsz, int(Lg))
            w = ops.pad_segments(wr, g_ptr, int(Lg)).view(bsz, int(Lg)) + is_pad * wp
        y = th.cat([p_vec, gv, gv - p_vec, gv * p_vec, pl, gl, pl_inv, gl_inv], dim=1)
        y = self.act(self.pred_fc1(y))
        y = self.pred_fc2(th.cat([y, pl, gl, pl_inv, gl_inv], dim=1))
        return y, w


class MeanPredictNet(PredictNet):
    ragged_pool = "mean"

    def agg_graph(self, g_rep, g_mask=None):
        return th.mean(g_rep, dim=1)


class SumPredictNet(PredictNet):
    ragged_pool = "sum"

    def agg_graph(self, g_rep, g_mask=None):
        return th.sum(g_rep, dim=1)


class MaxPredictNet(PredictNet):
    def agg_graph(self, g_rep, g_mask=None):
        return th.max(g_rep, dim=1)[0]
