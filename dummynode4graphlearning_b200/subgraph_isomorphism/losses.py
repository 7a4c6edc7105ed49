"""Back-propagated loss of the counting models: ``train_epoch`` in
subgraph_isomorphism/train.py:609-629 (criteria) and :776-813 (terms).

    bp_loss = crit(pred_c, counts)
            + rep_reg_w   * sum over {p_v, p_e, g_v, g_e}_rep of crit(rep, 0, slope 1) * rep.size(1)
            + match_loss_w * (crit(pred_v, node_weights) * Lg + crit(pred_e, edge_weights) * Le)
            + match_reg_w  * (crit(relu(pred_v - pred_c), 0, slope 0) * Lg + ... pred_e ...)

with crit(pred, target, slp) = {l1, mse, smooth_l1}(leaky_relu(pred, slp), target).  The match terms use the per-node /
per-edge weights of ``matching.node_weights`` / ``edge_weights`` left-padded to (B, L) (``pad_match_weights``), refined by
the model (``refine_node_weights``), and -- exactly as the reference does -- zero BOTH the targets and the predictions
on masked rows (padding, dummy nodes, reversed edges) in place and outside autograd before the criteria are applied.
Pure tensor code: it runs wherever the model's outputs live.
"""
import math

import torch as th
import torch.nn.functional as F

_CRITERIA = {"MAE": F.l1_loss, "MSE": F.mse_loss, "SMSE": F.smooth_l1_loss}


NUM_CYCLES = 2        # constants.py:39
_PI = 3.141592653589793


def scheduled_value(spec, step, total_steps, num_cycles=NUM_CYCLES):
    """value of a loss-weight / slope option at training step `step` (train.py:648-740).  `spec` is a number, or the
    CLI form ``anneal_<shape>$a$b`` / ``cyclical_<shape>$a$b`` with shape in {linear, cosine, constant, none}
    (config.py defaults: neg_pred_slp ``anneal_cosine$1.0$0.01``, match_reg_w ``anneal_cosine$0.01$0.0``): within each
    of `num_cycles` cycles over `total_steps` the value moves from a to b during the first half; an annealed value then
    stays at b, a cyclical one returns to a (linear) or follows the full cosine period (utils/anneal.py:12-49,
    utils/cyclical.py:12-46, always called with num_init_steps = 0)."""
    if isinstance(spec, (int, float)):
        return float(spec)
    if spec.startswith("anneal_"):
        kind, cyclical = spec[7:], False
    elif spec.startswith("cyclical_"):
        kind, cyclical = spec[9:], True
    else:
        raise ValueError(spec)
    kind, a, b = kind.rsplit("$", 3)
    a, b = float(a), float(b)
    if step > total_steps or not kind or kind in ("none", "constant"):
        return b
    progress = float(num_cycles * step) / max(1, total_steps) % 1
    if kind == "linear":
        if progress < 0.5:
            return float(a + (b - a) * (progress * 2))
        return float(b + (a - b) * (progress * 2 - 1)) if cyclical else b
    if kind == "cosine":
        if progress < 0.5 or cyclical:
            return float(a + (b - a) * (1 - math.cos(_PI * progress * 2)) / 2)
        return b
    raise NotImplementedError(kind)


def bp_criterion(kind):
    """train.py:620-627"""
    if kind not in _CRITERIA:
        raise NotImplementedError(kind)
    fn = _CRITERIA[kind]
    return lambda pred, target, neg_slp: fn(F.leaky_relu(pred, neg_slp), target)


def eval_criterion(kind):
    """train.py:609-618 (AUC needs scikit-learn on the host and is left to the caller)."""
    if kind not in _CRITERIA:
        raise NotImplementedError(kind)
    fn = _CRITERIA[kind]
    return lambda pred, target: fn(F.relu(pred), target)


def pad_match_weights(weights, seg_ptr, L):
    """flat per-node / per-edge integer weights (N,) of a batch -> (B, L) float, left-padded like every other per-graph
    tensor of the counting path (GraphAdjDataset.batchify, dataset.py:1604-1636)."""
    from .. import ops
    w = weights.to(th.float32).view(-1, 1).contiguous()
    return ops.pad_segments(w, seg_ptr, int(L)).view(-1, int(L))


def counting_bp_loss(output, counts, node_weights=None, edge_weights=None, model=None, bp_loss="MSE", neg_slp=0.01,
                     rep_reg_w=0.0, match_loss_w=0.0, match_reg_w=0.0):
    """-> (bp_loss, terms) with terms = dict(rep_reg, match_v_loss, match_e_loss, match_v_reg, match_e_reg) (detached
    scalars, the quantities the reference logs).  `counts` (B,) or (B, 1); node_weights (B, Lg) / edge_weights (B, Le)
    padded like output["g_v_mask"] / output["g_e_mask"], or None."""
    crit = bp_criterion(bp_loss)
    pred_c = output["pred_c"]
    counts = counts.to(pred_c.dtype).view(-1, 1)
    loss = crit(pred_c, counts, neg_slp)
    zero = pred_c.new_zeros(1)
    terms = dict(match_v_loss=zero, match_e_loss=zero, match_v_reg=zero, match_e_reg=zero)
    for kind, weights, pred_key, mask_key in (("v", node_weights, "pred_v", "g_v_mask"), ("e", edge_weights, "pred_e", "g_e_mask")):
        pred = output.get(pred_key)
        if weights is None or pred is None:
            continue
        with th.no_grad():
            weights = weights.to(pred.device).float()
            if model is not None:
                weights = (model.refine_node_weights if kind == "v" else model.refine_edge_weights)(weights)
            weights = weights.masked_fill(~output[mask_key], 0)
            pred.masked_fill_(~output[mask_key], 0)          # in place and outside autograd, as train.py:783-784
        L = pred.size(1)
        terms["match_%s_loss" % kind] = crit(pred, weights, neg_slp) * L
        terms["match_%s_reg" % kind] = crit(F.relu(pred - pred_c), th.zeros_like(pred), 0) * L
    rep_reg = zero
    for k in ("p_v_rep", "p_e_rep", "g_v_rep", "g_e_rep"):
        rep = output.get(k)
        if rep is not None:
            rep_reg = rep_reg + crit(rep, th.zeros_like(rep), 1) * rep.size(1)
    terms["rep_reg"] = rep_reg
    loss = loss + rep_reg_w * rep_reg
    loss = loss + match_loss_w * (terms["match_v_loss"] + terms["match_e_loss"])
    loss = loss + match_reg_w * (terms["match_v_reg"] + terms["match_e_reg"])
    return loss.view(()) if loss.numel() == 1 else loss, {k: v.detach() for k, v in terms.items()}
