"""Standalone CPU restatement of the reference's model math (TEST INFRASTRUCTURE ONLY).

Functional torch-CPU code, written edge-by-edge / graph-by-graph in the reference's own formulation
(per-edge gathered weights + bmm, Python-loop padding) so that it is an independent check of the
restructured CUDA path.  Parameters are read from a state_dict with the reference's key names, so the same
weights drive the reference classes (under oracle/shims), this restatement, and the product modules.

Pinned against the unmodified reference classes by tests/test_oracle_vs_reference.py (build container) and
by the golden fixtures under tests/golden/ (everywhere).  Each function cites the reference lines it follows.
"""
import math

import torch as th
import torch.nn.functional as F

LEAKY = 1 / 5.5  # subgraph_isomorphism/utils/act.py:27


def act_fn(name):
    return {
        "none": lambda x: x, "relu": F.relu, "relu6": F.relu6, "tanh": th.tanh, "sigmoid": th.sigmoid,
        "leaky_relu": lambda x: F.leaky_relu(x, LEAKY), "elu": F.elu, "gelu": F.gelu, "selu": F.selu, "celu": F.celu,
    }[name]


def scatter_sum(rows, index, n):
    """torch-scatter 2.0.7 CPU scatter_sum / DGL fn.sum: sequential adds in edge order."""
    out = th.zeros((n,) + tuple(rows.shape[1:]), dtype=rows.dtype)
    return out.index_add_(0, index, rows)


def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def batch_norm_train(sd, p, x, training=True):
    """nn.BatchNorm1d: training mode = biased batch statistics over all rows; eval mode = the running statistics."""
    if not training:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            training=False, eps=1e-5)
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, eps=1e-5)


def mlp_seq(sd, p, x, num_layers, act, batch_norm=False):
    """Linear [, BN], act, Linear ... as built at rgin.py:50-57 / dmpnn.py:45-60 (no act after the last)."""
    idx = 0
    for i in range(num_layers):
        x = linear(sd, "%s.%d" % (p, idx), x)
        idx += 1
        if i != num_layers - 1:
            if batch_norm:
                x = batch_norm_train(sd, "%s.%d" % (p, idx), x)
                idx += 1
            x = act(x)
            idx += 1
    return x


# ---------------------------------------------------------------------------------------------
# classification (graph_classification/graph_neural_networks/models)
def gin_mlp(sd, p, x, training=True):
    """Sequential(Linear, BN, ReLU, Linear, BN, ReLU)   gconv.py:190-196"""
    x = F.relu(batch_norm_train(sd, p + ".1", linear(sd, p + ".0", x), training))
    return F.relu(batch_norm_train(sd, p + ".4", linear(sd, p + ".3", x), training))


def gin_classifier(sd, x, edge_index, batch, num_graphs, num_layers, aggregation="sum", training=True):
    """GIN.forward, gconv.py:203-215, dropout 0.  GINConv [ext PyG 2.0.2]: nn(scatter_sum(x[src], dst) + (1+eps) x).
    training=False: model.eval() (BatchNorm uses its running statistics)."""
    src, dst = edge_index[0], edge_index[1]
    cnt = th.bincount(batch, minlength=num_graphs).clamp(min=1).float().view(-1, 1)

    def pool(h):
        s = scatter_sum(h, batch, num_graphs)
        return s if aggregation == "sum" else s / cnt

    out = 0
    for layer in range(num_layers):
        if layer == 0:
            x = gin_mlp(sd, "first_h", x, training)
            out = out + pool(linear(sd, "linears.0", x))
        else:
            agg = scatter_sum(x[src], dst, x.size(0))
            agg = agg + (1 + sd["convs.%d.eps" % (layer - 1)]) * x
            x = gin_mlp(sd, "nns.%d" % (layer - 1), agg, training)
            out = out + linear(sd, "linears.%d" % layer, pool(x))
    return F.log_softmax(out, dim=-1)


def rgin_classifier(sd, x, edge_index, edge_type, batch, num_graphs, num_layers, num_relations, aggregation="sum"):
    """RGIN.forward, rgconv.py:107-126; RGCNConv(aggr='add') [ext]: relations looped in order,
    out += scatter_sum(x[src_r], dst_r) @ W_r; + x @ root + bias."""
    src, dst = edge_index[0], edge_index[1]
    cnt = th.bincount(batch, minlength=num_graphs).clamp(min=1).float().view(-1, 1)

    def pool(h):
        s = scatter_sum(h, batch, num_graphs)
        return s if aggregation == "sum" else s / cnt

    out = 0
    for layer in range(num_layers):
        if layer == 0:
            x = gin_mlp(sd, "first_h", x)
            out = out + pool(linear(sd, "linears.0", x))
        else:
            p = "convs.%d" % (layer - 1)
            h = th.zeros(x.size(0), sd[p + ".weight"].size(2))
            for r in range(num_relations):
                sel = edge_type == r
                h = h + scatter_sum(x[src[sel]], dst[sel], x.size(0)) @ sd[p + ".weight"][r]
            h = h + x @ sd[p + ".root"] + sd[p + ".bias"]
            x = gin_mlp(sd, "nns.%d" % (layer - 1), h)
            out = out + linear(sd, "linears.%d" % layer, pool(x))
    return F.log_softmax(out, dim=-1)


# ---------------------------------------------------------------------------------------------
# counting layers (subgraph_isomorphism/models)
def rgin_layer(sd, p, h, src, dst, etype, cfg):
    """RGINLayer, rgin.py:102-160: per-edge weight gather + bmm, fn.sum, self loop, bias, mlp, act, act."""
    N, D = h.shape
    R, reg, nb = cfg["num_rels"], cfg["regularizer"], cfg["num_bases"]
    act = act_fn(cfg["act_func"])
    if reg in ("none", "basis"):
        w = sd[p + ".weight"]
        if (p + ".w_comp") in sd and sd[p + ".w_comp"] is not None:
            w = th.matmul(sd[p + ".w_comp"], w.view(w.size(0), -1)).view(R, D, -1)       # :103-106
        msg = th.bmm(h[src].unsqueeze(1), w.index_select(0, etype)).squeeze(1)          # :109-110
    else:
        si = D // nb
        w = sd[p + ".weight"].index_select(0, etype).view(-1, si, si)                    # :117
        msg = th.bmm(h[src].reshape(-1, 1, si), w).view(-1, D)                           # :118
    agg = scatter_sum(msg, dst, N)
    out = agg + h @ sd[p + ".loop_weight"] + sd[p + ".bias"]
    out = mlp_seq(sd, p + ".mlp", out, cfg["num_mlp_layers"], act, cfg.get("batch_norm", False))
    return act(act(out)) if cfg["num_mlp_layers"] == 0 else act(out)                     # :147-151


def rgcn_layer(sd, p, h, src, dst, etype, in_deg, out_deg, cfg):
    """RGCNLayer, rgcn.py:102-197: per-edge weight gather + bmm, per-edge norm (in: 1/(indeg[dst]+1); both:
    sqrt(outnorm[src] * innorm[dst]); :134-165), fn.sum, normalised self loop (:170-180), bias, [BatchNorm], act."""
    N, D = h.shape
    R, reg, nb, en = cfg["num_rels"], cfg["regularizer"], cfg["num_bases"], cfg["edge_norm"]
    act = act_fn(cfg["act_func"])
    if reg in ("none", "basis"):
        w = sd[p + ".weight"]
        if (p + ".w_comp") in sd and sd[p + ".w_comp"] is not None:
            w = th.matmul(sd[p + ".w_comp"], w.view(w.size(0), -1)).view(R, D, -1)       # :104-107
        msg = th.bmm(h[src].unsqueeze(1), w.index_select(0, etype)).squeeze(1)          # :110-111
    else:
        si = D // nb
        w = sd[p + ".weight"].index_select(0, etype).view(-1, si, si)                    # :120
        msg = th.bmm(h[src].reshape(-1, 1, si), w).view(-1, D)                           # :121
    innorm = (1.0 / (in_deg.to(h.dtype) + 1)).view(-1, 1)                               # self_loop is always on (:222-239)
    outnorm = (1.0 / (out_deg.to(h.dtype) + 1)).view(-1, 1)
    if en == "in":
        msg = msg * innorm[dst]
    elif en == "both":
        msg = msg * (outnorm[src] * innorm[dst]) ** 0.5
    agg = scatter_sum(msg, dst, N)
    loop = h @ sd[p + ".loop_weight"]
    if en == "in":
        loop = loop * innorm
    elif en == "both":
        loop = loop * (innorm * outnorm) ** 0.5
    out = agg + loop + sd[p + ".bias"]
    if cfg.get("batch_norm", False):
        out = batch_norm_train(sd, p + ".bn", out)
    return act(out)


def dmp_layer(sd, p, h, ef, src, dst, is_rev, out_deg, cfg):
    """DMPLayer, dmpnn.py:111-166."""
    act = act_fn(cfg["act_func"])
    W = {k: sd["%s.%s_weight" % (p, k)] for k in ("in", "out", "src", "dst", "nloop", "eloop")}
    hs, hd = h[src], h[dst]
    edge_msg = hd @ W["dst"] - hs @ W["src"]                                             # :112
    node_msg = -(ef @ W["in"])                                                           # :113
    if is_rev is not None:                                                               # :116-124
        r = is_rev.view(-1, 1)
        rev_edge_msg = hs @ W["dst"] - hd @ W["src"]
        rev_node_msg = ef @ W["out"]
        edge_msg = edge_msg.masked_fill(r, 0.0) + rev_edge_msg.masked_fill(~r, 0.0)
        node_msg = node_msg.masked_fill(r, 0.0) + rev_node_msg.masked_fill(~r, 0.0)
    agg = scatter_sum(node_msg, dst, h.size(0))
    n_out = h @ W["nloop"] + agg + sd[p + ".nbias"]                                      # :131-133
    n_out = mlp_seq(sd, p + ".nmlp", n_out, cfg["num_mlp_layers"], act, cfg.get("batch_norm", False))
    d = (1 + out_deg[dst].unsqueeze(-1).to(ef.dtype)).log2()                                 # :144-145
    add = 2 * (1 + d) * (ef @ (W["src"] - W["dst"]))                                     # :146
    e_out = ef @ W["eloop"] + add + edge_msg + sd[p + ".ebias"]                          # :147-149
    e_out = mlp_seq(sd, p + ".emlp", e_out, cfg["num_mlp_layers"], act, cfg.get("batch_norm", False))
    return n_out, e_out


def compgcn_layer(sd, p, h, ef, src, dst, is_rev, in_deg, out_deg, cfg):
    """CompGCNLayer, compgcn.py:169-279, in the reference's own per-edge formulation (self_loop is always on: the
    model never passes it, compgcn.py:311-321)."""
    act = act_fn(cfg["act_func"])
    opt, en = cfg["comp_opt"], cfg["edge_norm"]

    def comp(head, rel):                                                                 # :214-223
        if opt == "sub":
            return head - rel
        if opt == "mult":
            return head * rel
        n = head.size(-1)
        return th.fft.irfft(th.conj(th.fft.rfft(head, dim=-1)) * th.fft.rfft(rel.expand_as(head), dim=-1), n=n, dim=-1)

    innorm = (1.0 / (in_deg.to(h.dtype) + 1)).view(-1, 1)                               # :181-184
    outnorm = (1.0 / (out_deg.to(h.dtype) + 1)).view(-1, 1)                             # :191-194
    c = comp(h[src], ef)                                                                 # :226
    msg = c @ sd[p + ".in_weight"]
    if is_rev is not None:                                                               # :229-232
        r = is_rev.view(-1, 1)
        msg = msg.masked_fill(r, 0.0) + (c @ sd[p + ".out_weight"]).masked_fill(~r, 0.0)
    if en == "in":                                                                       # :203-208, :234-235
        msg = msg * innorm[dst]
    elif en == "out":
        msg = msg * outnorm[src]
    elif en == "both":
        msg = msg * (outnorm[src] * innorm[dst]) ** 0.5
    agg = scatter_sum(msg, dst, h.size(0))
    loop = comp(h, sd[p + ".loop_rel"]) @ sd[p + ".loop_weight"]                         # :245-248
    out = (agg + loop) * 0.3333333 + sd[p + ".bias"]
    if cfg.get("batch_norm", False):
        out = batch_norm_train(sd, p + ".bn", out)
    return act(out), ef @ sd[p + ".rel_weight"]                                          # :263-266


# ---------------------------------------------------------------------------------------------
# ragged <-> left-padded plumbing (utils/dl.py:51-127)
def pad_left(feats, sizes):
    """split_and_batchify_graph_feats(pre_pad=True): per-graph loop + cat."""
    L = int(max(sizes))
    rows, idx = [], 0
    for l in sizes:
        l = int(l)
        if l < L:
            rows.append(th.zeros((L - l,) + tuple(feats.shape[1:]), dtype=feats.dtype))
        rows.append(feats[idx: idx + l])
        idx += l
    return th.cat(rows, 0).view(len(sizes), L, -1)


def len_mask_left(sizes):
    """batch_convert_len_to_mask(pre_pad=True), dl.py:113-127 (for l == 0 the row stays all ones)."""
    L = int(max(sizes))
    m = th.ones((len(sizes), L), dtype=th.bool)
    for i, l in enumerate(sizes):
        l = int(l)
        if l > 0:
            m[i, : L - l] = False
    return m


def scalar_filter_gate(p_labels, p_sizes, g_labels, g_sizes):
    """get_filter_gate, basemodel.py:830-847 + ScalarFilter filter.py:10-16 on zero-left-padded label matrices."""
    p = pad_left(p_labels.view(-1, 1), p_sizes)   # B x Lp x 1
    g = pad_left(g_labels.view(-1, 1), g_sizes)   # B x Lg x 1
    gate = ((g.unsqueeze(2) - p.unsqueeze(1)) == 0).any(dim=2)   # B x Lg x 1
    Lg = g.size(1)
    out = [gate[i, Lg - int(l):] for i, l in enumerate(g_sizes)]
    return th.cat(out, 0).view(-1, 1).float()


def predict_net(sd, p, p_rep, p_mask, g_rep, g_mask, agg, act, with_weights):
    """PredictNet.forward, pred.py:87-156 (dropout 0).  agg pools over the WHOLE padded axis."""
    pool = {"sum": lambda t: t.sum(1), "mean": lambda t: t.mean(1), "max": lambda t: t.max(1)[0]}[agg]
    B, Lg = g_mask.shape
    pl = p_mask.to(p_rep.dtype).sum(1).view(B, 1)     # dtype of the parameters: the oracle also runs in float64
    gl = g_mask.to(p_rep.dtype).sum(1).view(B, 1)
    pli, gli = 1.0 / pl, 1.0 / gl
    pv = pool(linear(sd, p + ".p_fc", p_rep))
    g = linear(sd, p + ".g_fc", g_rep)
    w = None
    if with_weights:
        pe = pv.unsqueeze(1).expand(B, Lg, -1)
        plx, plix = pl.expand(B, Lg).unsqueeze(-1), pli.expand(B, Lg).unsqueeze(-1)
        w = act(linear(sd, p + ".weight_fc1", th.cat([pe, g, g - pe, g * pe, plx, plix], 2)))
        w = linear(sd, p + ".weight_fc2", th.cat([w, plx, plix], 2)).squeeze(-1)
    gv = pool(g)
    y = act(linear(sd, p + ".pred_fc1", th.cat([pv, gv, gv - pv, gv * pv, pl, gl, pli, gli], 1)))
    y = linear(sd, p + ".pred_fc2", th.cat([y, pl, gl, pli, gli], 1))
    return y, w


def _sizes(ptr):
    return [int(x) for x in (ptr[1:] - ptr[:-1])]


def _side(b):
    """int64 torch views of a flat batch dict."""
    t = lambda k: th.as_tensor(b[k]).long()
    s = dict(src=t("src"), dst=t("dst"), vid=t("vid"), vlabel=t("vlabel"), elabel=t("elabel"),
             n_sizes=_sizes(b["node_ptr"]), e_sizes=_sizes(b["edge_ptr"]), N=int(b["node_ptr"][-1]))
    for k in ("v_is_dummy", "e_is_dummy", "e_is_reversed"):
        s[k] = th.as_tensor(b[k]).bool() if k in b else None
    s["in_deg"] = th.bincount(s["dst"], minlength=s["N"])
    s["out_deg"] = th.bincount(s["src"], minlength=s["N"])
    return s


def counting_model(sd, pattern_b, graph_b, cfg):
    """GraphAdjModel.forward (basemodel.py:887-982, RGIN) / GraphAdjModelV2.forward (:1520-1703, DMPNN) with
    share_rep_net / share_enc_net as given by the state_dict keys.  cfg: dict(model='RGIN'|'DMPNN', hid_dim,
    num_layers, act_func, pred_act_func, pred_net, pred_with_enc, pred_with_deg, return_weights, filter,
    residual, add_node_id, node_pred, edge_pred, layer={...})."""
    P, G = _side(pattern_b), _side(graph_b)
    v2 = cfg["model"] in ("DMPNN", "CompGCN")
    act_pred = act_fn(cfg["pred_act_func"])
    agg = {"SumPredictNet": "sum", "MeanPredictNet": "mean", "MaxPredictNet": "max"}[cfg["pred_net"]]
    residual = cfg.get("residual", True)

    def enc(side, S):
        e = {"v": sd[side + "_enc_net.v.weight"][S["vid"]], "vl": sd[side + "_enc_net.vl.weight"][S["vlabel"]]}
        if v2:
            e["el"] = sd[side + "_enc_net.el.weight"][S["elabel"]]
        return e

    def emb(side, e):
        v = e["vl"] @ sd[side + "_emb_net.vl.weight"]
        if cfg.get("add_node_id", False):
            v = v + e["v"] @ sd[side + "_emb_net.v.weight"]
        if not v2:
            return v, None
        return v, e["el"] @ sd[side + "_emb_net.el.weight"]

    v_gate = e_gate = None
    if cfg.get("filter", True):
        v_gate = scalar_filter_gate(P["vlabel"], P["n_sizes"], G["vlabel"], G["n_sizes"])
        if v2:
            e_gate = scalar_filter_gate(P["elabel"], P["e_sizes"], G["elabel"], G["e_sizes"])

    def rep(side, S, v, e, vg, eg):
        if vg is not None:
            v = v * vg
        if eg is not None and e is not None:
            e = e * eg
        for i in range(cfg["num_layers"]):
            if not v2:
                if cfg["model"] == "RGCN":
                    p = "%s_rep_net.rgcn.%s_rgcn_(%d)" % (side, cfg["rep_name"][side], i)
                    o = rgcn_layer(sd, p, v, S["src"], S["dst"], S["elabel"], S["in_deg"], S["out_deg"], cfg["layer"][side])
                else:
                    p = "%s_rep_net.rgin.%s_rgin_(%d)" % (side, cfg["rep_name"][side], i)
                    o = rgin_layer(sd, p, v, S["src"], S["dst"], S["elabel"], cfg["layer"][side])
                if vg is not None:
                    o = o * vg
                v = v + o if residual and v.shape == o.shape else o
            else:
                if cfg["model"] == "CompGCN":
                    p = "%s_rep_net.compgcn.%s_compgcn_(%d)" % (side, cfg["rep_name"][side], i)
                    nv, ne = compgcn_layer(sd, p, v, e, S["src"], S["dst"], S["e_is_reversed"], S["in_deg"], S["out_deg"],
                                           cfg["layer"][side])
                else:
                    p = "%s_rep_net.dmpnn.%s_dmpnn_(%d)" % (side, cfg["rep_name"][side], i)
                    nv, ne = dmp_layer(sd, p, v, e, S["src"], S["dst"], S["e_is_reversed"], S["out_deg"], cfg["layer"][side])
                if vg is not None:
                    nv = nv * vg
                if eg is not None:
                    ne = ne * eg
                if residual and nv.shape == v.shape and ne.shape == e.shape:
                    v, e = v + nv, e + ne
                else:
                    v, e = nv, ne
        return v, e

    p_enc, g_enc = enc("p", P), enc("g", G)
    p_v_emb, p_e_emb = emb("p", p_enc)
    g_v_emb, g_e_emb = emb("g", g_enc)
    p_v_rep, p_e_rep = rep("p", P, p_v_emb, p_e_emb, None, None)
    g_v_rep, g_e_rep = rep("g", G, g_v_emb, g_e_emb, v_gate, e_gate)

    def node_out(S, e, r):
        feats = []
        if cfg["pred_with_enc"]:
            feats += [e["v"], e["vl"]]
        if cfg["pred_with_deg"]:
            feats += [S["out_deg"].float().view(-1, 1), S["in_deg"].float().view(-1, 1)]
        o = th.cat(feats + [r], -1)
        mask = len_mask_left(S["n_sizes"])
        if S["v_is_dummy"] is not None:
            mask = mask & ~pad_left(S["v_is_dummy"].view(-1, 1), S["n_sizes"]).view(mask.shape)
        o = pad_left(o, S["n_sizes"]).masked_fill(~mask.unsqueeze(-1), 0)
        return o, mask

    def edge_out(S, e, r):
        u, v = S["src"], S["dst"]
        feats = []
        if cfg["pred_with_enc"]:
            feats += [e["v"][u], e["v"][v], e["vl"][u], e["el"], e["vl"][v]]
        if cfg["pred_with_deg"]:
            feats += [S["out_deg"].float().view(-1, 1)[u], S["in_deg"].float().view(-1, 1)[v]]
        o = th.cat(feats + [r], -1)
        mask = len_mask_left(S["e_sizes"])
        for k in ("e_is_dummy", "e_is_reversed"):
            if S[k] is not None:
                mask = mask & ~pad_left(S[k].view(-1, 1), S["e_sizes"]).view(mask.shape)
        o = pad_left(o, S["e_sizes"]).masked_fill(~mask.unsqueeze(-1), 0)
        return o, mask

    out = dict(p_v_emb=p_v_emb, p_e_emb=p_e_emb, g_v_emb=g_v_emb, g_e_emb=g_e_emb, p_v_rep=p_v_rep, p_e_rep=p_e_rep,
               g_v_rep=g_v_rep, g_e_rep=g_e_rep)
    rw = cfg.get("return_weights", "none")
    if not v2:
        po, pm = node_out(P, p_enc, p_v_rep)
        go, gm = node_out(G, g_enc, g_v_rep)
        y, w = predict_net(sd, "pred_net", po, pm, go, gm, agg, act_pred, "node" in rw)
        out.update(pred_c=y, pred_v=w, pred_e=None, p_v_mask=pm, g_v_mask=gm, p_e_mask=None, g_e_mask=None)
        return out
    po, pm = node_out(P, p_enc, p_v_rep)
    go, gm = node_out(G, g_enc, g_v_rep)
    peo, pem = edge_out(P, p_enc, p_e_rep)
    geo, gem = edge_out(G, g_enc, g_e_rep)
    vc = vw = ec = ew = None
    if cfg.get("node_pred", True):
        vc, vw = predict_net(sd, "pred_net.v", po, pm, go, gm, agg, act_pred, "node" in rw)
    if cfg.get("edge_pred", True):
        ec, ew = predict_net(sd, "pred_net.e", peo, pem, geo, gem, agg, act_pred, "edge" in rw)
    if vc is not None and ec is not None:                                   # basemodel.py:1506-1512
        gvl, gel = gm.to(vc.dtype).sum(1).view(-1, 1), gem.to(vc.dtype).sum(1).view(-1, 1)
        y = (gvl / (gvl + gel)) * vc + (gel / (gvl + gel)) * ec
    else:
        y = vc if vc is not None else ec
    out.update(pred_c=y, pred_v=vw, pred_e=ew, p_v_mask=pm, g_v_mask=gm, p_e_mask=pem, g_e_mask=gem)
    return out


def counting_loss(out, counts, rep_reg_w=0.0, neg_slp=0.01):
    """bp_loss of train_epoch with bp_loss='MSE' (train.py:624-625, 776-813), match terms off."""
    crit = lambda pred, target, slp: F.mse_loss(F.leaky_relu(pred, slp), target)
    loss = crit(out["pred_c"], counts.to(out["pred_c"].dtype).view(-1, 1), neg_slp)
    reg = 0.0
    for k in ("p_v_rep", "p_e_rep", "g_v_rep", "g_e_rep"):
        if out.get(k) is not None:
            reg = reg + crit(out[k], th.zeros_like(out[k]), 1) * out[k].size(1)
    return loss + rep_reg_w * reg


def counting_cfg_from_kwargs(name, kw):
    """oracle.models.counting_model config equivalent to the constructor kwargs `kw`."""
    shared = kw.get("share_rep_net", True)

    pre = "rep_rgcn" if name == "RGCN" else "rep_rgin"

    def nb(n_rel):
        b = kw.get(pre + "_num_bases", -1)
        reg = kw.get(pre + "_regularizer", "basis")
        return n_rel if (reg == "none" or b is None or b > n_rel or b <= 0) else b

    n_layers = kw.get("rep_num_graph_layers", 1)
    if name == "RGIN":
        lay = dict(num_rels=kw["max_ngel"], regularizer=kw.get("rep_rgin_regularizer", "basis"), num_bases=nb(kw["max_ngel"]),
                   num_mlp_layers=kw.get("rep_rgin_num_mlp_layers", 2), act_func=kw.get("rep_act_func", "relu"),
                   batch_norm=kw.get("rep_rgin_batch_norm", False))
        layp = dict(lay) if shared else dict(lay, num_rels=kw["max_npel"], num_bases=nb(kw["max_npel"]))
    elif name == "RGCN":
        lay = dict(num_rels=kw["max_ngel"], regularizer=kw.get("rep_rgcn_regularizer", "basis"), num_bases=nb(kw["max_ngel"]),
                   edge_norm=kw.get("rep_rgcn_edge_norm", "in"), act_func=kw.get("rep_act_func", "relu"),
                   batch_norm=kw.get("rep_rgcn_batch_norm", False))
        layp = dict(lay) if shared else dict(lay, num_rels=kw["max_npel"], num_bases=nb(kw["max_npel"]))
    elif name == "CompGCN":
        lay = dict(comp_opt=kw.get("rep_compgcn_comp_opt", "mult"), edge_norm=kw.get("rep_compgcn_edge_norm", "none"),
                   act_func=kw.get("rep_act_func", "relu"), batch_norm=kw.get("rep_compgcn_batch_norm", False))
        layp = dict(lay)
    else:
        lay = dict(num_mlp_layers=kw.get("rep_dmpnn_num_mlp_layers", 2), act_func=kw.get("rep_act_func", "relu"),
                   batch_norm=kw.get("rep_dmpnn_batch_norm", False))
        layp = dict(lay)
    return dict(model=name, num_layers=n_layers, pred_act_func=kw.get("pred_act_func", "relu"),
                pred_net=kw.get("pred_net", "SumPredictNet"), pred_with_enc=kw.get("pred_with_enc", False),
                pred_with_deg=kw.get("pred_with_deg", False), return_weights=kw.get("pred_return_weights", "none"),
                filter=kw.get("filter_net", "None") == "ScalarFilter", residual=kw.get("rep_residual", True),
                add_node_id=kw.get("add_node_id", False), node_pred=kw.get("node_pred", True),
                edge_pred=kw.get("edge_pred", True), rep_name={"g": "graph", "p": "graph" if shared else "pattern"},
                layer={"g": lay, "p": layp})
