"""shared test plumbing: golden loading, oracle configuration from constructor kwargs, comparisons."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def oracle_cfg(name, kw):
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


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def assert_close_rel(a, b, tol, what="", atol=0.0):
    """max-abs error normalised by the max-abs reference value (the survey's parity measure).  `atol` is an
    absolute floor for quantities that are mathematically zero (e.g. the gradient of a Linear bias feeding a
    BatchNorm), where only rounding noise is left on both sides."""
    assert tuple(a.shape) == tuple(b.shape), (what, a.shape, b.shape)
    if b.numel() == 0:
        return
    if atol > 0 and float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max()) <= atol:
        return
    e = rel_err(a, b)
    assert e <= tol, "%s: relative error %.3e > %.1e" % (what, e, tol)


def batches_equal(a, b, keys):
    for k in keys:
        x = a[k].cpu().numpy() if isinstance(a[k], torch.Tensor) else np.asarray(a[k])
        y = b[k].cpu().numpy() if isinstance(b[k], torch.Tensor) else np.asarray(b[k])
        assert x.shape == y.shape, (k, x.shape, y.shape)
        assert np.array_equal(x, y), (k, np.flatnonzero(x != y)[:8])


def oracle_preprocess_chain(p, g, cfg, remove_loops, add_rev, add_dummy, convert_conj):
    """the data-set level switches in the order of train.py:1271-1340, each with the maxima its predecessors leave."""
    npe, npel, nge, ngel = cfg["max_npe"], cfg["max_npel"], cfg["max_nge"], cfg["max_ngel"]
    if remove_loops:
        p, g = _OT().sub_remove_loops(p), _OT().sub_remove_loops(g)
    if add_rev:
        p, g = _OT().sub_add_reversed(p, npe, npel), _OT().sub_add_reversed(g, nge, ngel)
        npe, npel, nge, ngel = 2 * npe, 2 * npel, 2 * nge, 2 * ngel
    if add_dummy:
        p = _OT().sub_add_dummy(p, cfg["max_npv"], cfg["max_npvl"], npe, npel)
        g = _OT().sub_add_dummy(g, cfg["max_ngv"], cfg["max_ngvl"], nge, ngel)
    if convert_conj:
        p, g = _OT().sub_conjugate(p), _OT().sub_conjugate(g)
    return p, g


def _OT():
    from oracle import transforms
    return transforms
