"""Live pin of the oracle: executes the UNMODIFIED reference from /root/reference under oracle/shims on fresh seeds
(not the committed golden ones) and compares with the standalone restatement (oracle/transforms.py, oracle/models.py).
Build container only -- skipped wherever /root/reference does not exist (e.g. the GPU box)."""
import zlib

import numpy as np
import pytest
import torch

from dummynode4graphlearning_b200 import synth
from dummynode4graphlearning_b200.transforms import process_model_config
from helpers import (NASTY_CFG, assert_close_rel, batches_equal, nasty_sub_batch, nasty_tu_batch, oracle_cfg,
                     oracle_preprocess_chain)
from oracle import models as OM
from oracle import transforms as OT

pytestmark = pytest.mark.reference_live

TU_KEYS = ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "vid", "eid", "v_is_dummy", "e_is_dummy")
SUB_KEYS = ("node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "v_is_dummy", "eid", "elabel", "e_is_dummy", "e_is_reversed")
CONJ_KEYS = ("node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "v_is_dummy", "v_is_reversed", "eid", "elabel", "e_is_dummy")


@pytest.mark.parametrize("shape,nb,seed", [("mutag", 9, 101), ("proteins", 5, 102), ("mutag", 3, 103)])
def test_tu_transforms_live(shape, nb, seed):
    from oracle import ref_drive as rd
    b = synth.tu_batch(shape, nb, seed=seed)
    ref_dummy = rd.ref_tu_load(b, True)
    batches_equal(OT.tu_add_dummy(b), rd.igraphs_to_batch(ref_dummy), [k for k in TU_KEYS if k in ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "v_is_dummy", "e_is_dummy")])
    ref_conj = rd.igraphs_to_batch(rd.ref_tu_conjugate(ref_dummy))
    got = OT.tu_conjugate(OT.tu_add_dummy(b))
    batches_equal(got, ref_conj, [k for k in TU_KEYS if k in ref_conj and k in got])
    ref_line = rd.igraphs_to_batch(rd.ref_tu_conjugate(rd.ref_tu_load(b, False)))
    got = OT.tu_conjugate(b)
    batches_equal(got, ref_line, [k for k in ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel") if k in ref_line])


@pytest.mark.parametrize("shape,bs,seed", [("small", 5, 111), ("small", 3, 112)])
def test_sub_transforms_live(shape, bs, seed):
    from oracle import ref_drive as rd
    p, g, _ = synth.counting_batch(shape, bs, seed=seed)
    cfg = synth.counting_config(shape)
    rp, rg = rd.ref_sub_add_dummy(p, g, cfg)
    pd_ = OT.sub_add_dummy(p, cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"])
    gd_ = OT.sub_add_dummy(g, cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
    batches_equal(pd_, rp, SUB_KEYS)
    batches_equal(gd_, rg, SUB_KEYS)
    batches_equal(OT.sub_conjugate(pd_), rd.ref_sub_conjugate(rp), CONJ_KEYS)
    batches_equal(OT.sub_conjugate(gd_), rd.ref_sub_conjugate(rg), CONJ_KEYS)
    c = dict(cfg, add_rev=False, add_dummy=True, convert_conj=True)
    assert process_model_config(c) == rd.ref_process_model_config(c)


@pytest.mark.parametrize("tag,name,over", [
    ("live/RGIN", "RGIN", dict(hid_dim=16, pred_hid_dim=16)),
    ("live/RGCN", "RGCN", dict(hid_dim=16, pred_hid_dim=16, rep_rgcn_edge_norm="both")),
    ("live/CompGCN", "CompGCN", dict(hid_dim=16, pred_hid_dim=16, rep_compgcn_comp_opt="sub", rep_compgcn_edge_norm="both")),
    ("live/DMPNN", "DMPNN", dict(hid_dim=16, pred_hid_dim=16, node_pred=True, edge_pred=True, pred_return_weights="node,edge")),
    ("live/RGIN_position", "RGIN", dict(hid_dim=16, pred_hid_dim=16, enc_net="Position")),            # --enc_net Position
    ("live/DMPNN_position", "DMPNN", dict(hid_dim=16, pred_hid_dim=16, enc_net="Position", node_pred=True, edge_pred=True)),
])
def test_counting_models_live(tag, name, over):
    """the reference's own RGIN / DMPNN classes (fake-DGL graph drives their message / update UDFs) vs the functional
    restatement: forward tensors, loss and every parameter gradient."""
    import torch.nn.functional as F
    from oracle import ref_drive as rd
    p, g, counts = synth.counting_batch("small", 6, seed=121)
    cfg = dict(synth.counting_config("small"), add_dummy=True)
    mc = process_model_config(cfg)
    pd_ = OT.sub_add_dummy(p, cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"])
    gd_ = OT.sub_add_dummy(g, cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
    kw = rd.counting_kwargs({k: v for k, v in mc.items() if k.startswith("max_")}, **over)
    model = rd.ref_counting_model(name, kw, seed=zlib.crc32(tag.encode()) % 1000)
    o = model(rd.dgl_batched(pd_), rd.dgl_batched(gd_))
    c = torch.from_numpy(counts).float().view(-1, 1)
    loss = F.mse_loss(F.leaky_relu(o["pred_c"], 0.01), c)
    loss.backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    out = OM.counting_model(sd, pd_, gd_, oracle_cfg(name, kw))
    for k in ("pred_c", "p_v_rep", "g_v_rep", "g_e_rep"):
        if o[k] is not None:
            assert_close_rel(out[k], o[k].detach(), 1e-6, k)
    mine = OM.counting_loss(out, torch.from_numpy(counts), rep_reg_w=0.0)
    assert_close_rel(mine, loss.detach(), 1e-6, "loss")
    mine.backward()
    for n, pp in model.named_parameters():
        if pp.grad is None:
            continue
        got = sd[n].grad
        alias = n.replace("g_rep_net", "p_rep_net", 1) if n.startswith("g_rep_net") else None
        if alias in sd and sd[alias].grad is not None:
            got = got + sd[alias].grad if got is not None else sd[alias].grad
        assert_close_rel(got, pp.grad, 1e-5, "grad " + n)


@pytest.mark.parametrize("shape,bs,seed", [("small", 5, 301), ("large", 1, 302)])
def test_augmentation_flags_live(shape, bs, seed):
    """oracle vs the reference's remove_loops / add_reversed_edges / norms / eigenvalues, fresh seeds."""
    from oracle import ref_drive as rd
    p, g, _ = synth.counting_batch(shape, bs, seed=seed)
    cfg = synth.counting_config(shape)
    g2 = dict(g)
    g2["dst"] = g["dst"].copy()
    g2["dst"][::4] = g2["src"][::4]
    _, rg = rd.ref_sub_remove_loops(p, g2)
    batches_equal(OT.sub_remove_loops(g2), rg, ("edge_ptr", "src", "dst", "eid", "elabel"))
    rp, rg = rd.ref_sub_add_reversed(p, g, cfg)
    og = OT.sub_add_reversed(g, cfg["max_nge"], cfg["max_ngel"])
    batches_equal(og, rg, ("edge_ptr", "src", "dst", "eid", "elabel", "e_is_reversed"))
    batches_equal(OT.sub_add_reversed(p, cfg["max_npe"], cfg["max_npel"]), rp, ("edge_ptr", "src", "dst", "eid", "elabel", "e_is_reversed"))
    for sl in (True, False):
        r = rd.ref_sub_norms_eigen(og, sl)
        nn_, en = OT.compute_norm(og, sl)
        np.testing.assert_array_equal(nn_, r["node_norm"])
        np.testing.assert_array_equal(en, r["edge_norm"])
    ne, ee = OT.compute_largest_eigenvalues(og)
    np.testing.assert_array_equal(np.repeat(np.maximum(ne, 1), np.diff(og["node_ptr"])), r["node_eigenv"].ravel())
    np.testing.assert_array_equal(np.repeat(np.maximum(ee, 1), np.diff(og["edge_ptr"])), r["edge_eigenv"].ravel())


@pytest.mark.parametrize("shape,nb,seed", [("mutag", 7, 401), ("proteins", 4, 402)])
def test_tu_io_live(shape, nb, seed):
    """load_tu_dir == the reference's load_graph_data_from_TUDatadir on the same files, and
    save_tu_dir's text == the reference's save_graph_data text for the reference's own CONJ graphs."""
    import tempfile
    from dummynode4graphlearning_b200.graph_classification import io as tuio
    from oracle import ref_drive as rd
    b = synth.tu_batch(shape, nb, seed=seed)
    with tempfile.TemporaryDirectory() as d:
        raw = rd.write_tu_files(b, d)                 # 0-based labels on disk, "u, v" with a blank: what the parser must accept
        mine = tuio.load_tu_dir(raw)
    ref = rd.igraphs_to_batch(rd.ref_tu_load(b, False))
    batches_equal(mine, ref, ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel"))
    if "vattr" in ref:
        np.testing.assert_array_equal(mine["vattr"], ref["vattr"])
    conj_graphs = rd.ref_tu_conjugate(rd.ref_tu_load(b, True))
    files = rd.ref_tu_save(conj_graphs)
    lines = tuio.tu_file_lines(rd.igraphs_to_batch(conj_graphs))
    for suffix, ref_lines in files.items():
        assert lines[suffix] == ref_lines, suffix


@pytest.mark.parametrize("shape,bs,seed", [("small", 10, 501), ("large", 2, 503)])
def test_match_weights_live(shape, bs, seed):
    """oracle vs the reference's numba loops on fresh seeds."""
    from oracle import ref_drive as rd
    p, g, _ = synth.counting_batch(shape, bs, seed=seed)
    mats = synth.random_subisomorphisms(p, g, seed=seed)
    rn, re = rd.ref_match_weights(mats, p, g)
    np.testing.assert_array_equal(OT.subiso_node_weights(mats, g), rn)
    np.testing.assert_array_equal(OT.subiso_edge_weights(mats, p, g), re)
    for a, r in zip(OT.subiso_conjugate(mats, p, g), rd.ref_conjugate_subisomorphisms(mats, p, g)):
        assert a.shape == r.shape
        np.testing.assert_array_equal(a, r)
    assert rn.sum() > 0 and (shape != "small" or re.sum() > 0)


@pytest.mark.parametrize("name", ["RGIN", "DMPNN"])
def test_position_encoder_matches_reference_class(name):
    """--enc_net Position (basemodel.py:642-646, embed.py:211-222): the product's sinusoid tables are bit-identical to
    the reference class's and its models expose the same state_dict keys and shapes."""
    from dummynode4graphlearning_b200.subgraph_isomorphism import models as PM
    from oracle import refload
    ns = refload.subgraph()
    for d, n in ((14, 65), (10, 17), (10, 18), (2, 2)):
        assert torch.equal(PM.PositionEmbedding(d, n).weight, ns.embed.PositionEmbedding(d, n).weight)
    mc = process_model_config(dict(synth.counting_config("small"), add_dummy=True))
    kw = dict({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=16, enc_net="Position", emb_net="Equivariant",
              filter_net="ScalarFilter", rep_num_graph_layers=2, rep_num_pattern_layers=2, share_enc_net=False)
    mine = getattr(PM, name)(**kw).state_dict()
    ref = {"RGIN": ns.rgin.RGIN, "DMPNN": ns.dmpnn.DMPNN}[name](**kw).state_dict()
    assert sorted(mine) == sorted(ref)
    for k in mine:
        assert mine[k].shape == ref[k].shape, k
        if "enc_net" in k:
            assert torch.equal(mine[k], ref[k]) and not mine[k].requires_grad, k


def test_activation_registry_matches_reference():
    """every --rep_act_func / --pred_act_func choice (utils/act.py:457-473) exists with the reference's values and
    gradients, including the row-wise ones (sparsemax, maximum, minimum) on ties, 3-D inputs and other dims."""
    from dummynode4graphlearning_b200.subgraph_isomorphism import utils as PU
    from oracle import refload
    import importlib
    refload.subgraph()                       # installs the shims and puts the reference's package root on sys.path
    ref = importlib.import_module("utils.act")
    assert sorted(PU.supported_act_funcs) == sorted(ref.supported_act_funcs)
    g = torch.Generator().manual_seed(11)
    xs = [torch.randn(7, 16, generator=g), torch.randn(3, 5, 8, generator=g) * 4,
          torch.tensor([[1.0, 1.0, 0.5, -2.0], [0.0, 0.0, 0.0, 0.0], [-1.0, -3.0, -1.0, -3.0]])]
    for name in sorted(ref.supported_act_funcs):
        if name in ("gumbel_softmax", "prelu"):
            continue
        for x in xs:
            a = x.clone().requires_grad_(True)
            b = x.clone().requires_grad_(True)
            ya, yb = PU.map_activation_str_to_layer(name)(a), ref.map_activation_str_to_layer(name)(b)
            assert torch.equal(ya, yb), name
            w = torch.arange(ya.numel(), dtype=ya.dtype).view(ya.shape).cos()
            (ya * w).sum().backward()
            (yb * w).sum().backward()
            assert torch.allclose(a.grad, b.grad, rtol=0, atol=1e-7), name
    for cls in ("Sparsemax", "Maximum", "Minimum"):
        x = torch.randn(4, 6, 5, generator=g)
        for dim in (0, 1, -1):
            kw = [dict(dim=dim)] if cls == "Sparsemax" else [dict(dim=dim), dict(dim=dim, scale_up=True)]
            for k in kw:
                assert torch.allclose(getattr(PU, cls)(**k)(x), getattr(ref, cls)(**k)(x), rtol=0, atol=1e-7), (cls, k)
    y = PU.Maximum(inplace=True, scale_up=True)
    z1, z2 = x.clone(), x.clone()
    assert torch.equal(y(z1), ref.Maximum(inplace=True, scale_up=True)(z2)) and torch.equal(z1, z2)
    torch.manual_seed(5)
    ya = PU.map_activation_str_to_layer("gumbel_softmax")(xs[0])
    torch.manual_seed(5)
    yb = ref.map_activation_str_to_layer("gumbel_softmax")(xs[0])
    assert torch.equal(ya, yb)
    assert abs(PU.supported_act_funcs["prelu"].weight.item() - ref.supported_act_funcs["prelu"].weight.item()) < 1e-7


@pytest.mark.parametrize("name,extra", [("RGIN", {}), ("DMPNN", {}), ("CompGCN", {}),
                                        ("RGIN", dict(share_emb_net=False, share_enc_net=False, pred_with_enc=False))])
def test_model_expand_matches_reference(name, extra):
    """BaseModel.expand (basemodel.py:167-219, called by train.py:1399 on every loaded checkpoint): after loading the
    reference's weights and expanding both models to larger maxima with the same RNG state, the state_dicts are
    identical (trailing-corner copies, zeros elsewhere, fresh layers for new heads, aliasing quirk included)."""
    from dummynode4graphlearning_b200.subgraph_isomorphism import models as PM
    from oracle import refload
    ns = refload.subgraph()
    small = dict(max_ngv=33, max_ngvl=9, max_nge=200, max_ngel=10, max_npv=9, max_npvl=9, max_npe=24, max_npel=10)
    big = dict(max_ngv=65, max_ngvl=17, max_nge=384, max_ngel=10, max_npv=17, max_npvl=17, max_npe=48, max_npel=10)
    common = dict(hid_dim=16, emb_net="Equivariant", filter_net="ScalarFilter", rep_num_graph_layers=2,
                  rep_num_pattern_layers=2, pred_with_enc=True, pred_with_deg=True, pred_hid_dim=16)
    kw = dict(small, **common)
    kw.update(extra)
    torch.manual_seed(0)
    ref = getattr(ns.models, name)(**kw)
    mine = getattr(PM, name)(**kw)
    mine.load_state_dict(ref.state_dict())
    kw2 = dict(big, **common)
    kw2.update(extra, pred_return_weights="node")
    torch.manual_seed(1)
    ref.expand(**kw2)
    torch.manual_seed(1)
    mine.expand(**kw2)
    sa, sb = mine.state_dict(), ref.state_dict()
    assert sorted(sa) == sorted(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert (mine.p_emb_net is mine.g_emb_net) == (ref.p_emb_net is ref.g_emb_net)
    assert all(getattr(mine, k) == getattr(ref, k) for k in big)
    with pytest.raises(ValueError):
        mine.expand(base=3)


@pytest.mark.parametrize("name,kind,seed", [("DMPNN", "SMSE", 7), ("RGIN", "MAE", 8), ("CompGCN", "MSE", 9)])
def test_bp_loss_live_against_reference_train_epoch(name, kind, seed):
    """one mini-batch through the reference's own train_epoch (verbatim, with its match terms, criteria
    and clipping) vs the product's losses.counting_bp_loss applied to the same reference model's outputs: loss and every
    parameter gradient."""
    from dummynode4graphlearning_b200.subgraph_isomorphism.losses import counting_bp_loss
    from oracle import ref_drive as rd
    p, g, counts = synth.counting_batch("small", 5, seed=seed)
    cfg = dict(synth.counting_config("small"), add_dummy=True)
    mc = process_model_config(cfg)
    pd_ = OT.sub_add_dummy(p, cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"])
    gd_ = OT.sub_add_dummy(g, cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
    over = dict(hid_dim=16, pred_hid_dim=16, pred_return_weights="node,edge" if name != "RGIN" else "node")
    if name != "RGIN":
        over.update(node_pred=True, edge_pred=True)
    kw = rd.counting_kwargs({k: v for k, v in mc.items() if k.startswith("max_")}, **over)
    model = rd.ref_counting_model(name, kw, seed=seed)
    with torch.no_grad():
        for n, q in model.named_parameters():
            if "weight_fc2" in n:
                q.normal_(0.0, 0.05)
            elif "pred_fc2" in n:
                q.normal_(0.0, 0.02)
    gen = torch.Generator().manual_seed(seed)
    nw = torch.randint(0, 4, (5, int(np.diff(gd_["node_ptr"]).max())), generator=gen)
    ew = torch.randint(0, 4, (5, int(np.diff(gd_["edge_ptr"]).max())), generator=gen)
    conf = dict(bp_loss=kind, neg_pred_slp=0.05, match_loss_w=0.3, match_reg_w=0.2, rep_reg_w=1e-3, max_grad_norm=0.0)
    _, ref_loss, ref_grads = rd.ref_train_epoch(model, rd.dgl_batched(pd_), rd.dgl_batched(gd_), counts, nw.clone(),
                                                ew.clone(), conf)
    model.zero_grad()
    out = model(rd.dgl_batched(pd_), rd.dgl_batched(gd_))
    loss, terms = counting_bp_loss(out, torch.from_numpy(counts), nw, ew, model=model, bp_loss=kind, neg_slp=0.05,
                                   rep_reg_w=1e-3, match_loss_w=0.3, match_reg_w=0.2)
    assert abs(float(loss.detach()) - ref_loss) <= 1e-6 * abs(ref_loss)
    assert float(terms["match_v_loss"]) > 0 and float(terms["rep_reg"]) > 0
    loss.backward()
    for n, q in model.named_parameters():
        if ref_grads[n] is None:
            assert q.grad is None, n
        else:
            assert_close_rel(q.grad, ref_grads[n], 1e-6, "grad " + n)


@pytest.mark.parametrize("flags", [(True, True, True, True), (False, True, True, False), (True, False, True, True),
                                   (False, True, False, True)])
def test_preprocessing_chain_live(flags):
    """remove_loops -> add_reversed_edges -> add_dummy_nodes_edges -> convert_to_conjugate chained as train.py's main
    does (:1271-1340), reference functions vs the oracle chain that CountingPipeline.augment mirrors on the GPU."""
    from oracle import ref_drive as rd
    remove_loops, add_rev, add_dummy, convert_conj = flags
    p, g, _ = synth.counting_batch("small", 7, seed=sum(flags) * 31 + 5)
    g = dict(g)
    g["dst"] = g["dst"].copy()
    g["dst"][::5] = g["src"][::5]                     # plant loops
    cfg = synth.counting_config("small")
    rp, rg = p, g
    npe, npel, nge, ngel = cfg["max_npe"], cfg["max_npel"], cfg["max_nge"], cfg["max_ngel"]
    if remove_loops:
        rp, rg = rd.ref_sub_remove_loops(rp, rg)
    if add_rev:
        rp, rg = rd.ref_sub_add_reversed(rp, rg, cfg)
        npe, npel, nge, ngel = 2 * npe, 2 * npel, 2 * nge, 2 * ngel
    if add_dummy:
        rp, rg = rd.ref_sub_add_dummy(rp, rg, dict(cfg, max_npe=npe, max_npel=npel, max_nge=nge, max_ngel=ngel))
    if convert_conj:
        rp, rg = rd.ref_sub_conjugate(rp), rd.ref_sub_conjugate(rg)
    op, og = oracle_preprocess_chain(p, g, cfg, *flags)
    keys = [k for k in ("node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "eid", "elabel", "v_is_dummy", "e_is_dummy",
                        "e_is_reversed", "v_is_reversed") if k in rg and k in og]
    assert {"node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "eid", "elabel"} <= set(keys)
    batches_equal(op, rp, keys)
    batches_equal(og, rg, keys)


def test_loss_weight_schedules_match_reference():
    """neg_pred_slp / match_loss_w / match_reg_w / rep_reg_w option strings resolved per step exactly as train_epoch does
    (train.py:648-740 -> utils/anneal.py, utils/cyclical.py with num_init_steps = 0)."""
    import importlib
    from dummynode4graphlearning_b200.subgraph_isomorphism.losses import scheduled_value
    from oracle import refload
    refload.subgraph()
    anneal_fn = importlib.import_module("utils.anneal").anneal_fn
    cyclical_fn = importlib.import_module("utils.cyclical").cyclical_fn
    total = 1000
    for spec in ("anneal_cosine$1.0$0.01", "anneal_cosine$0.01$0.0", "anneal_linear$0.5$2.0", "cyclical_cosine$0.0$1.0",
                 "cyclical_linear$1.0$0.25", "anneal_constant$3.0$4.0", "cyclical_none$3.0$4.0"):
        head, a, b = spec.rsplit("$", 3)
        for step in list(range(0, 1003, 7)) + [249, 250, 251, 499, 500, 501, 999, 1000, 1001, 5000]:
            if head.startswith("anneal_"):
                ref = anneal_fn(head[7:], step, num_init_steps=0, num_anneal_steps=total, num_cycles=2, value1=float(a),
                                value2=float(b))
            else:
                ref = cyclical_fn(head[9:], step, num_init_steps=0, num_cyclical_steps=total, num_cycles=2,
                                  value1=float(a), value2=float(b))
            assert scheduled_value(spec, step, total) == ref, (spec, step)
    assert scheduled_value(0.25, 3, 10) == 0.25 and scheduled_value(1, 3, 10) == 1.0
    with pytest.raises(ValueError):
        scheduled_value("cosine$1$2", 0, 10)


@pytest.mark.parametrize("seed", range(12))
def test_tu_transforms_fuzz_live(seed):
    """dummy augmentation, CONJ and LINE transforms on graphs with self loops, duplicate edges, isolated and single nodes and
    edgeless graphs: oracle == reference."""
    from oracle import ref_drive as rd
    rng = np.random.default_rng(seed)
    b = nasty_tu_batch(rng, int(rng.integers(1, 6)))
    ref_dummy = rd.ref_tu_load(b, True)
    batches_equal(OT.tu_add_dummy(b), rd.igraphs_to_batch(ref_dummy),
                  ["node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "v_is_dummy", "e_is_dummy"])
    ref_conj = rd.igraphs_to_batch(rd.ref_tu_conjugate(ref_dummy))
    got = OT.tu_conjugate(OT.tu_add_dummy(b))
    batches_equal(got, ref_conj, [k for k in TU_KEYS if k in ref_conj and k in got])
    ref_line = rd.igraphs_to_batch(rd.ref_tu_conjugate(rd.ref_tu_load(b, False)))
    batches_equal(OT.tu_conjugate(b), ref_line, [k for k in ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel") if k in ref_line])


@pytest.mark.parametrize("seed", range(12))
def test_sub_preprocessing_fuzz_live(seed):
    """random subsets of remove_loops / add_reversed_edges / add_dummy_nodes_edges / convert_to_conjugate chained in
    train.py's order on multigraphs with loops, isolated nodes and edgeless graphs: oracle == reference."""
    from oracle import ref_drive as rd
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 5))
    p, g = nasty_sub_batch(rng, B, 5, 4), nasty_sub_batch(rng, B, 8, 6)
    flags = tuple(bool(x) for x in rng.integers(0, 2, 4))
    cfg = NASTY_CFG
    rp, rg = p, g
    npe, npel, nge, ngel = cfg["max_npe"], cfg["max_npel"], cfg["max_nge"], cfg["max_ngel"]
    if flags[0]:
        rp, rg = rd.ref_sub_remove_loops(rp, rg)
    if flags[1]:
        rp, rg = rd.ref_sub_add_reversed(rp, rg, cfg)
        npe, npel, nge, ngel = 2 * npe, 2 * npel, 2 * nge, 2 * ngel
    if flags[2]:
        rp, rg = rd.ref_sub_add_dummy(rp, rg, dict(cfg, max_npe=npe, max_npel=npel, max_nge=nge, max_ngel=ngel))
    if flags[3]:
        rp, rg = rd.ref_sub_conjugate(rp), rd.ref_sub_conjugate(rg)
    op, og = oracle_preprocess_chain(p, g, cfg, *flags)
    keys = [k for k in ("node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "eid", "elabel", "v_is_dummy", "e_is_dummy",
                        "e_is_reversed", "v_is_reversed") if k in rg and k in og]
    batches_equal(op, rp, keys)
    batches_equal(og, rg, keys)


@pytest.mark.parametrize("name,over", [("RGIN", {}), ("DMPNN", dict(node_pred=True, edge_pred=True)),
                                       ("RGCN", dict(rep_rgcn_edge_norm="both")),
                                       ("CompGCN", dict(rep_compgcn_comp_opt="sub", rep_compgcn_edge_norm="both"))])
def test_counting_models_fuzz_live(name, over):
    """the four counting models on dummy-augmented multigraphs with self loops, repeated edges and isolated nodes
    (every graph keeps at least one non-loop edge: with none, the reference's own 1 / (number of real edges) is inf):
    oracle == reference for pred_c and every parameter gradient."""
    import torch.nn.functional as F
    from oracle import ref_drive as rd
    cfg = dict(NASTY_CFG, add_dummy=True)
    mc = process_model_config(cfg)
    done = 0
    for seed in range(2000, 2120):
        rng = np.random.default_rng(seed)
        B = int(rng.integers(1, 5))
        p, g = nasty_sub_batch(rng, B, 5, 4), nasty_sub_batch(rng, B, 8, 6)
        if any((b["src"][a:z] == b["dst"][a:z]).all() for b in (p, g) for a, z in zip(b["edge_ptr"][:-1], b["edge_ptr"][1:])):
            continue                                   # an edgeless (or loops-only) graph: degenerate for the reference itself
        counts = rng.poisson(3.0, B).astype(np.int64)
        pd_ = OT.sub_add_dummy(p, cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"])
        gd_ = OT.sub_add_dummy(g, cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
        kw = rd.counting_kwargs({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=16, pred_hid_dim=16, **over)
        model = rd.ref_counting_model(name, kw, seed=seed)
        with torch.no_grad():
            for n, q in model.named_parameters():
                if "fc2" in n:
                    q.normal_(0.0, 0.1)
        o = model(rd.dgl_batched(pd_), rd.dgl_batched(gd_))
        F.mse_loss(F.leaky_relu(o["pred_c"], 0.01), torch.from_numpy(counts).float().view(-1, 1)).backward()
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
        out = OM.counting_model(sd, pd_, gd_, oracle_cfg(name, kw))
        assert_close_rel(out["pred_c"], o["pred_c"].detach(), 1e-5, "pred_c seed %d" % seed)
        OM.counting_loss(out, torch.from_numpy(counts), rep_reg_w=0.0).backward()
        for n, pp in model.named_parameters():
            if pp.grad is None:
                continue
            got = sd[n].grad
            alias = n.replace("g_rep_net", "p_rep_net", 1) if n.startswith("g_rep_net") else None
            if alias in sd and sd[alias].grad is not None:
                got = got + sd[alias].grad if got is not None else sd[alias].grad
            assert_close_rel(got, pp.grad, 1e-4, "grad %s seed %d" % (n, seed))
        done += 1
        if done == 5:
            break
    assert done == 5


def test_match_weights_fuzz_live():
    """match weights and the conjugate subisomorphism map on multigraphs with self loops and repeated (u, v, label)
    edges (every pattern / graph keeps at least one edge: the reference's numba loops take max() of the edge arrays)."""
    from oracle import ref_drive as rd
    done = 0
    for seed in range(3000, 3200):
        rng = np.random.default_rng(seed)
        B = int(rng.integers(1, 5))
        p, g = nasty_sub_batch(rng, B, 5, 4), nasty_sub_batch(rng, B, 8, 6)
        if (np.diff(p["edge_ptr"]) == 0).any() or (np.diff(g["edge_ptr"]) == 0).any():
            continue
        mats = synth.random_subisomorphisms(p, g, seed=seed, max_rows=6)
        rn, re = rd.ref_match_weights(mats, p, g)
        np.testing.assert_array_equal(OT.subiso_node_weights(mats, g), rn)
        np.testing.assert_array_equal(OT.subiso_edge_weights(mats, p, g), re)
        for a, r in zip(OT.subiso_conjugate(mats, p, g), rd.ref_conjugate_subisomorphisms(mats, p, g)):
            assert a.shape == r.shape
            np.testing.assert_array_equal(a, r)
        done += 1
        if done == 15:
            break
    assert done == 15


@pytest.mark.parametrize("seed", range(6))
def test_tu_io_fuzz_live(seed):
    """reader and writer on the nasty TU batches: load_tu_dir == the reference's loader, and the text written for the
    reference's own DUMMY_ / CONJ_ graphs == save_graph_data's."""
    import tempfile
    from dummynode4graphlearning_b200.graph_classification import io as tuio
    from oracle import ref_drive as rd
    rng = np.random.default_rng(500 + seed)
    b = nasty_tu_batch(rng, int(rng.integers(1, 6)))
    with tempfile.TemporaryDirectory() as d:
        mine = tuio.load_tu_dir(rd.write_tu_files(b, d))
    ref = rd.igraphs_to_batch(rd.ref_tu_load(b, False))
    batches_equal(mine, ref, ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel"))
    dummy_graphs = rd.ref_tu_load(b, True)
    for graphs in (dummy_graphs, rd.ref_tu_conjugate(dummy_graphs)):
        files = rd.ref_tu_save(graphs)
        lines = tuio.tu_file_lines(rd.igraphs_to_batch(graphs))
        for suffix, ref_lines in files.items():
            assert lines[suffix] == ref_lines, suffix


@pytest.mark.parametrize("name,add,variant,seed", [
    ("GIN", {"train_eps": True, "num_layers": 3, "aggregation": "sum"}, "conj", 0),
    ("GIN", {"train_eps": False, "num_layers": 2, "aggregation": "mean"}, "dummy", 1),
    ("GIN", {"train_eps": True, "num_layers": 2, "aggregation": "sum"}, "line", 2),
    ("RGIN", {"num_layers": 2}, "dummy", 3),
    ("RGIN", {"num_layers": 3}, "conj", 4),
])
def test_classification_models_fuzz_live(name, add, variant, seed):
    """the reference's own GIN / classification RGIN classes (gconv.py, rgconv.py, run under oracle/shims) on the PyG view
    of DUMMY_ / CONJ_ / LINE_ multigraphs with self loops, repeated edges, isolated nodes and edgeless graphs (the view
    removes loops and merges repeats, so nodes without any edge and graphs of a single vertex reach the layers):
    log-softmax outputs, loss and every gradient of the oracle restatement == the reference."""
    import torch.nn.functional as F
    from argparse import Namespace
    from oracle import ref_drive as rd
    for attempt in range(50):     # LINE_ of an edgeless graph has no vertices (nothing a loader would keep): draw again
        rng = np.random.default_rng(7000 + seed + 100 * attempt)
        raw = nasty_tu_batch(rng, int(rng.integers(3, 7)))
        b = {"dummy": lambda: OT.tu_add_dummy(raw), "conj": lambda: OT.tu_conjugate(OT.tu_add_dummy(raw)),
             "line": lambda: OT.tu_conjugate(raw)}[variant]()
        if int(np.diff(b["node_ptr"]).min()) > 0:
            break
    else:
        raise AssertionError("no usable batch")
    vl, el = np.asarray(b["vlabel"], np.int64), np.asarray(b["elabel"], np.int64)
    vmin, emin = int(vl.min()), int(el.min()) if len(el) else 0
    R = (int(el.max()) - emin + 1) if len(el) else 1
    s, d, first, mult = OT.pyg_coalesce(b["src"], b["dst"], el - emin, R)
    x = torch.from_numpy(np.eye(int(vl.max()) - vmin + 1, dtype=np.float32)[vl - vmin])
    edge_index = torch.from_numpy(np.stack([s, d]).astype(np.int64))
    edge_attr = torch.from_numpy(mult.astype(np.float32))
    B = int(b["num_graphs"])
    batch = torch.from_numpy(np.repeat(np.arange(B), np.diff(b["node_ptr"])).astype(np.int64))
    y = torch.from_numpy(np.asarray(raw["y"], np.int64))
    args = Namespace(num_features=x.size(1), hidden_dim=8, nhid=8, num_classes=2, dropout_ratio=0.0, additional=add,
                     epochs=1, device="cpu", num_relations=R)
    model = rd.ref_classifier(name, args, seed=seed)
    with torch.no_grad():
        for n, p_ in model.named_parameters():
            if n.endswith("eps"):
                p_.fill_(0.2)
    model.train()
    out_ref = model(Namespace(x=x, edge_index=edge_index, edge_attr=edge_attr, batch=batch, y=y))
    loss_ref = F.nll_loss(out_ref, y)
    loss_ref.backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    nl = add["num_layers"]
    if name == "GIN":
        out = OM.gin_classifier(sd, x, edge_index, batch, B, nl, add["aggregation"])
    else:
        et = edge_attr.max(1)[1] if edge_attr.numel() else torch.zeros(0, dtype=torch.int64)
        out = OM.rgin_classifier(sd, x, edge_index, et, batch, B, nl, R)
    assert_close_rel(out, out_ref.detach(), 1e-6, "log_softmax")
    loss = F.nll_loss(out, y)
    assert_close_rel(loss, loss_ref.detach(), 1e-6, "loss")
    loss.backward()
    for n, p_ in model.named_parameters():
        if p_.grad is None:
            continue
        got = sd[n].grad
        if got is None and n.startswith("convs.") and ".nn." in n:     # aliases of nns.* (gconv.py:195-197)
            got = sd[n.replace("convs.", "nns.").replace(".nn.", ".")].grad
        assert_close_rel(got, p_.grad, 2e-5, "grad " + n, atol=2e-5)
    if name == "GIN":     # model.eval(): BatchNorm on the running statistics the training pass above has just updated
        model.eval()
        with torch.no_grad():
            out_eval_ref = model(Namespace(x=x, edge_index=edge_index, edge_attr=edge_attr, batch=batch, y=y))
        sd_eval = {k: v.detach().clone() for k, v in model.state_dict().items()}
        out_eval = OM.gin_classifier(sd_eval, x, edge_index, batch, B, nl, add["aggregation"], training=False)
        assert_close_rel(out_eval, out_eval_ref, 1e-6, "eval-mode log_softmax")
