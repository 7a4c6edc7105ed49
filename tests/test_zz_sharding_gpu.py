"""GPU: exact single-process numbers from a sharded counting mini-batch (SURVEY.md 8(e) i, v).

The counting head sums over the PADDED axis (padded rows contribute the bias, App. A-7) and the label filter sees the
zero padding of patterns shorter than the batch's longest one (App. A-14), so a rank that only holds a shard must pad
to the batch-wide maxima: ``BatchedGraph.set_padded_lengths`` (fed by ``parallel.sync_padded_lengths`` in
``CountingPipeline(exact_sharding=True)``).  Here the two shards are evaluated in one process."""
import numpy as np
import pytest
import torch

from helpers import assert_close_rel

pytestmark = pytest.mark.gpu


def _slice(b, lo, hi):
    """samples [lo, hi) of a flat block-diagonal batch dict (host numpy)."""
    n0, n1, e0, e1 = int(b["node_ptr"][lo]), int(b["node_ptr"][hi]), int(b["edge_ptr"][lo]), int(b["edge_ptr"][hi])
    out = dict(num_graphs=hi - lo, node_ptr=(b["node_ptr"][lo:hi + 1] - n0).astype(np.int32),
               edge_ptr=(b["edge_ptr"][lo:hi + 1] - e0).astype(np.int32),
               src=(b["src"][e0:e1] - n0).astype(np.int32), dst=(b["dst"][e0:e1] - n0).astype(np.int32))
    for k in ("vid", "vlabel"):
        out[k] = b[k][n0:n1]
    for k in ("eid", "elabel"):
        out[k] = b[k][e0:e1]
    return out


def _build(name, device, over):
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.subgraph_isomorphism.models import DMPNN, RGIN
    cfg = dict(synth.counting_config("small"), add_dummy=True)
    mc = T.process_model_config(cfg)
    kw = dict({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=64, rep_num_graph_layers=2,
              rep_num_pattern_layers=2, rep_act_func="leaky_relu", pred_act_func="leaky_relu", pred_net="SumPredictNet",
              pred_hid_dim=64, emb_net="Equivariant", enc_net="Multihot", filter_net="ScalarFilter", pred_with_enc=True,
              pred_with_deg=True, rep_rgin_regularizer="bdd", rep_rgin_num_bases=4, init_neigenv=4.0, init_eeigenv=4.0)
    kw.update(over)
    torch.manual_seed(3)
    model = {"RGIN": RGIN, "DMPNN": DMPNN}[name](**kw)
    with torch.no_grad():
        for n, q in model.named_parameters():
            # pred_fc2 / weight_fc2 are zero-initialised in the reference (App. A-8); biases are randomised so that the
            # bias of padded rows (App. A-7) is visible in pred_c
            if "pred_fc2" in n or "weight_fc2" in n or (n.endswith("bias") and "pred_net" in n):
                q.normal_(0.0, 0.1)
    return model.to(device).train(), cfg      # dropout 0, no BatchNorm: train mode is batch-independent here


def _graphs(p, g, cfg, device):
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.graph import BatchedGraph
    pd_ = T.sub_add_dummy(T.to_device(p, device), cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"])
    gd_ = T.sub_add_dummy(T.to_device(g, device), cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
    return BatchedGraph.from_batch(pd_, device), BatchedGraph.from_batch(gd_, device)


@pytest.mark.parametrize("name,over", [("RGIN", {}), ("DMPNN", dict(node_pred=True, edge_pred=True))])
def test_sharded_batch_with_batch_wide_padding_equals_full_batch(device, name, over):
    from dummynode4graphlearning_b200 import synth
    B, cut = 24, 9
    p, g, _ = synth.counting_batch("small", B, seed=21)
    model, cfg = _build(name, device, over)
    with torch.no_grad():
        pattern, graph = _graphs(p, g, cfg, device)
        full = model(pattern, graph)["pred_c"].clone()
        Lg_v, Lg_e = graph.max_num_nodes(), graph.max_num_edges()
        Lp_v, Lp_e = pattern.max_num_nodes(), pattern.max_num_edges()
        parts, plain = [], []
        for lo, hi in ((0, cut), (cut, B)):
            ps, gs = _graphs(_slice(p, lo, hi), _slice(g, lo, hi), cfg, device)
            plain.append(model(ps, gs)["pred_c"].clone())
            ps, gs = _graphs(_slice(p, lo, hi), _slice(g, lo, hi), cfg, device)
            assert gs.padded_num_nodes() == gs.max_num_nodes()
            gs.set_padded_lengths(Lg_v, Lg_e)
            ps.set_padded_lengths(Lp_v, Lp_e)
            assert gs.padded_num_nodes() == Lg_v and ps.padded_num_edges() == Lp_e
            parts.append(model(ps, gs)["pred_c"].clone())
    assert_close_rel(torch.cat(parts), full, 1e-5, "sharded pred_c with batch-wide padding")
    # and the override matters: at least one shard pads shorter on its own, which changes the head's bias term
    sizes = np.diff(g["node_ptr"])
    if sizes[:cut].max() != sizes[cut:].max():
        assert float((torch.cat(plain) - full).abs().max()) > 1e-6 * float(full.abs().max())


def test_padded_length_override_rejects_too_short(device):
    from dummynode4graphlearning_b200 import synth
    p, g, _ = synth.counting_batch("small", 4, seed=2)
    _, cfg = _build("RGIN", device, {})
    _, graph = _graphs(p, g, cfg, device)
    with pytest.raises(ValueError):
        graph.set_padded_lengths(graph.max_num_nodes() - 1, None)
    graph.set_padded_lengths(None, graph.max_num_edges() + 3)
    assert graph.padded_num_edges() == graph.max_num_edges() + 3 and graph.padded_num_nodes() == graph.max_num_nodes()
