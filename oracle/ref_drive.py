"""Drive the UNMODIFIED reference functions on batches in this repo's flat layout.

TEST SCAFFOLDING; needs /root/reference (build container only).  Converts a batch dict
(dummynode4graphlearning_b200/synth.py layout) into the containers the reference expects
(TU text files / fake igraph / fake DGL graphs), calls the reference's own code, converts
the result back.  Used by oracle/gen_golden.py and tests/test_oracle_vs_reference.py.
"""
import os
import tempfile

import numpy as np
import torch as th

from . import refload
from .shims import fake_dgl


# ------------------------------------------------------------------------------------------
# classification flavour (tu_data_processing.py)
def write_tu_files(b, root, name="T"):
    """TU text layout (what tu_data_processing.py:134-152 parses)."""
    raw = os.path.join(root, name)
    os.makedirs(raw, exist_ok=True)
    assert int(b["vlabel"].min()) == 1 and int(b["elabel"].min()) == 1, "labels must already have min 1"

    def w(suffix, rows):
        with open(os.path.join(raw, "%s_%s.txt" % (name, suffix)), "w") as f:
            for r in rows:
                f.write(r + "\n")

    w("A", ["%d, %d" % (s + 1, d + 1) for s, d in zip(b["src"], b["dst"])])
    gi = np.repeat(np.arange(b["num_graphs"]) + 1, np.diff(b["node_ptr"]))
    w("graph_indicator", [str(int(x)) for x in gi])
    w("node_labels", [str(int(x) - 1) for x in b["vlabel"]])          # 0-based on disk -> +1 (line 158-159)
    if b.get("has_edge_labels", True):
        w("edge_labels", [str(int(x) - 1) for x in b["elabel"]])
    if "vattr" in b:
        w("node_attributes", [repr(float(x)) for x in b["vattr"]])
    w("graph_labels", [str(int(x)) for x in b.get("y", np.zeros(b["num_graphs"], int))])
    return raw


def igraphs_to_batch(graphs):
    B = len(graphs)
    node_ptr = np.zeros(B + 1, np.int32)
    edge_ptr = np.zeros(B + 1, np.int32)
    src, dst = [], []
    cols_v, cols_e = {}, {}
    for i, g in enumerate(graphs):
        node_ptr[i + 1] = node_ptr[i] + g.vcount()
        edge_ptr[i + 1] = edge_ptr[i] + g.ecount()
        el = g.get_edgelist()
        src += [a + int(node_ptr[i]) for a, _ in el]
        dst += [c + int(node_ptr[i]) for _, c in el]
        for k in g.vertex_attributes():
            cols_v.setdefault(k, []).extend(g.vs[k])
        for k in g.edge_attributes():
            cols_e.setdefault(k, []).extend(g.es[k])
    o = dict(num_graphs=B, node_ptr=node_ptr, edge_ptr=edge_ptr,
             src=np.asarray(src, np.int32), dst=np.asarray(dst, np.int32))
    names = {"LABEL": "label", "IS_DUMMY": "_is_dummy", "ID": "id", "ATTR": "attr"}
    for k, v in cols_v.items():
        key = "v" + names[k] if k != "IS_DUMMY" else "v_is_dummy"
        o[key] = np.asarray(v, np.float32 if k == "ATTR" else np.int32)
    for k, v in cols_e.items():
        key = "e" + names[k] if k != "IS_DUMMY" else "e_is_dummy"
        o[key] = np.asarray(v, np.float32 if k == "ATTR" else np.int32)
    return o


def ref_tu_load(b, with_dummy):
    """reference load_graph_data_from_TUDatadir (tu_data_processing.py:125-220) on batch b."""
    tu = refload.classification().tu
    with tempfile.TemporaryDirectory() as d:
        raw = write_tu_files(b, d)
        return tu.load_graph_data_from_TUDatadir(raw, with_dummy=with_dummy)


def ref_tu_conjugate(graphs):
    tu = refload.classification().tu
    return [tu.convert_conjugate_graph_forward(g) for g in graphs]


def ref_tu_save(graphs):
    """reference save_graph_data (tu_data_processing.py:353-414) -> dict suffix -> lines."""
    tu = refload.classification().tu
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "X", "raw")
        os.makedirs(out)
        tu.save_graph_data(graphs, out)
        res = {}
        for fn in sorted(os.listdir(out)):
            with open(os.path.join(out, fn)) as f:
                res[fn[len("X_"):-4]] = [ln.strip() for ln in f]
        return res


# ------------------------------------------------------------------------------------------
# subgraph-isomorphism flavour (train.py / utils/graph.py) on fake DGL graphs
def batch_to_dgl_list(b):
    """per-graph fake DGLGraphs with ndata{id,label[,is_dummy]} / edata{id,label[,..]} (int64)."""
    C = refload.subgraph().constants
    gs = []
    for g in range(b["num_graphs"]):
        n0, n1 = int(b["node_ptr"][g]), int(b["node_ptr"][g + 1])
        e0, e1 = int(b["edge_ptr"][g]), int(b["edge_ptr"][g + 1])
        G = fake_dgl.DGLGraph(b["src"][e0:e1].astype(np.int64) - n0, b["dst"][e0:e1].astype(np.int64) - n0, n1 - n0)
        G.ndata[C.NODEID] = th.from_numpy(b["vid"][n0:n1].astype(np.int64))
        G.ndata[C.NODELABEL] = th.from_numpy(b["vlabel"][n0:n1].astype(np.int64))
        G.edata[C.EDGEID] = th.from_numpy(b["eid"][e0:e1].astype(np.int64))
        G.edata[C.EDGELABEL] = th.from_numpy(b["elabel"][e0:e1].astype(np.int64))
        for k, fr, (a, z) in (("v_is_dummy", G.ndata, (n0, n1)), ("e_is_dummy", G.edata, (e0, e1))):
            if k in b:
                fr[C.DUMMYFLAG] = th.from_numpy(b[k][a:z].astype(bool))
        if "e_is_reversed" in b:
            G.edata[C.REVFLAG] = th.from_numpy(b["e_is_reversed"][e0:e1].astype(bool))
        if "v_is_reversed" in b:
            G.ndata[C.REVFLAG] = th.from_numpy(b["v_is_reversed"][n0:n1].astype(bool))
        gs.append(G)
    return gs


def dgl_list_to_batch(gs):
    C = refload.subgraph().constants
    B = len(gs)
    node_ptr = np.zeros(B + 1, np.int32)
    edge_ptr = np.zeros(B + 1, np.int32)
    for i, g in enumerate(gs):
        node_ptr[i + 1] = node_ptr[i] + g.number_of_nodes()
        edge_ptr[i + 1] = edge_ptr[i] + g.number_of_edges()
    o = dict(num_graphs=B, node_ptr=node_ptr, edge_ptr=edge_ptr)
    o["src"] = np.concatenate([g._u.numpy() + node_ptr[i] for i, g in enumerate(gs)]).astype(np.int32)
    o["dst"] = np.concatenate([g._v.numpy() + node_ptr[i] for i, g in enumerate(gs)]).astype(np.int32)
    nmap = {C.NODEID: "vid", C.NODELABEL: "vlabel", C.DUMMYFLAG: "v_is_dummy", C.REVFLAG: "v_is_reversed",
            C.INDEGREE: "in_deg", C.OUTDEGREE: "out_deg"}
    emap = {C.EDGEID: "eid", C.EDGELABEL: "elabel", C.DUMMYFLAG: "e_is_dummy", C.REVFLAG: "e_is_reversed"}
    for k, name in nmap.items():
        if all(k in g.ndata for g in gs):
            o[name] = np.concatenate([g.ndata[k].numpy() for g in gs]).astype(np.int32)
    for k, name in emap.items():
        if all(k in g.edata for g in gs):
            o[name] = np.concatenate([g.edata[k].numpy() for g in gs]).astype(np.int32)
    return o


def ref_sub_add_dummy(pattern_b, graph_b, cfg):
    """reference add_dummy_nodes_edges, GraphAdj branch (train.py:404-474), called with the
    dataset maxima exactly as train.py:1322-1334 does."""
    tf = refload.subgraph().train_funcs
    ps, gs = batch_to_dgl_list(pattern_b), batch_to_dgl_list(graph_b)
    ds = tf.GraphAdjDataset(
        [{"pattern": p, "graph": g, "counts": 0, "subisomorphisms": th.zeros((0, 0), dtype=th.long)}
         for p, g in zip(ps, gs)])
    tf.add_dummy_nodes_edges(ds, cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"],
                             cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
    return dgl_list_to_batch([x["pattern"] for x in ds]), dgl_list_to_batch([x["graph"] for x in ds])


def _ref_dataset(pattern_b, graph_b):
    tf = refload.subgraph().train_funcs
    ps, gs = batch_to_dgl_list(pattern_b), batch_to_dgl_list(graph_b)
    return tf, tf.GraphAdjDataset(
        [{"pattern": p, "graph": g, "counts": 0, "subisomorphisms": th.zeros((0, 0), dtype=th.long)}
         for p, g in zip(ps, gs)])


def ref_sub_add_reversed(pattern_b, graph_b, cfg):
    """reference add_reversed_edges, GraphAdj branch (train.py:291-345), called as train.py:1315 does."""
    tf, ds = _ref_dataset(pattern_b, graph_b)
    tf.add_reversed_edges(ds, cfg["max_npe"], cfg["max_npel"], cfg["max_nge"], cfg["max_ngel"])
    return dgl_list_to_batch([x["pattern"] for x in ds]), dgl_list_to_batch([x["graph"] for x in ds])


def ref_sub_remove_loops(pattern_b, graph_b):
    """reference remove_loops, GraphAdj branch (train.py:270-288)."""
    tf, ds = _ref_dataset(pattern_b, graph_b)
    tf.remove_loops(ds)
    return dgl_list_to_batch([x["pattern"] for x in ds]), dgl_list_to_batch([x["graph"] for x in ds])


def ref_sub_norms_eigen(graph_b, self_loop=True):
    """reference calculate_norms + calculate_eigenvalues (train.py:500-527) on every graph of the batch."""
    C = refload.subgraph().constants
    tf, ds = _ref_dataset(graph_b, graph_b)
    tf.calculate_norms(ds, self_loop=self_loop)
    tf.calculate_eigenvalues(ds)
    gs = [x["graph"] for x in ds]
    return dict(node_norm=th.cat([g.ndata[C.NORM] for g in gs]).numpy(), edge_norm=th.cat([g.edata[C.NORM] for g in gs]).numpy(),
                node_eigenv=th.cat([g.ndata[C.NODEEIGENV] for g in gs]).numpy(),
                edge_eigenv=th.cat([g.edata[C.EDGEEIGENV] for g in gs]).numpy())


def ref_match_weights(mats, pattern_b, graph_b):
    """the reference's numba loops (dataset.py:54-108) driven per sample exactly as calculate_node_weights /
    calculate_edge_weights do (dataset.py:1491-1520): pattern edges in eid order, graph edges in srcdst order, result
    scattered back through g_eid."""
    df = refload.subgraph().dataset_funcs
    nw, ew = [], []
    for b, (P, G) in enumerate(zip(batch_to_dgl_list(pattern_b), batch_to_dgl_list(graph_b))):
        m = np.asarray(mats[b], dtype=np.int64)
        if m.size == 0:
            nw.append(np.zeros(G.number_of_nodes(), np.int64))
            ew.append(np.zeros(G.number_of_edges(), np.int64))
            continue
        nw.append(df.compute_nodeseq_subisoweights(G.number_of_nodes(), m))
        C = refload.subgraph().constants
        p_u, p_v, p_e = P.all_edges(form="all", order="eid")
        p_el = P.edata[C.EDGELABEL][p_e]
        g_u, g_v, g_e = G.all_edges(form="all", order="srcdst")
        g_el = G.edata[C.EDGELABEL][g_e]
        w = th.zeros(g_e.size(0), dtype=th.long)
        w[g_e] = th.from_numpy(df.compute_edgeseq_subisoweights(p_u.numpy(), p_v.numpy(), p_el.numpy(), g_u.numpy(),
                                                                g_v.numpy(), g_el.numpy(), m))
        ew.append(w.numpy())
    return np.concatenate(nw), np.concatenate(ew)


def ref_conjugate_subisomorphisms(mats, pattern_b, graph_b):
    """the reference's get_conjugate_subisomorphisms (utils/graph.py:294-330, numba) driven per sample exactly as
    convert_to_conjugate's GraphAdj branch does (train.py:569-587, without its final ``mask``, which is a no-op for the
    non-negative ids it produces)."""
    gu_ = refload.subgraph().graph_utils
    C = refload.subgraph().constants
    out = []
    for b, (P, G) in enumerate(zip(batch_to_dgl_list(pattern_b), batch_to_dgl_list(graph_b))):
        m = np.asarray(mats[b], dtype=np.int64)
        if m.size == 0 or P.number_of_edges() == 0:
            out.append(np.zeros((0, P.number_of_edges()), np.int64))
            continue
        p_u, p_v, p_e = P.all_edges(form="all", order="eid")
        p_el = P.edata[C.EDGELABEL][p_e]
        g_u, g_v, g_e = G.all_edges(form="all", order="srcdst")
        g_el = G.edata[C.EDGELABEL][g_e]
        cs = gu_.get_conjugate_subisomorphisms(p_u.numpy(), p_v.numpy(), p_el.numpy(), g_u.numpy(), g_v.numpy(), g_el.numpy(), m)
        out.append(np.vstack([g_e.numpy()[cs[i]] for i in range(len(cs))]))
    return out


def write_counting_dirs(root, pattern_b, graph_b, counts, mats, layout="own", creator="igraph version 0.9.11"):
    """dump a synthetic counting batch in the directory layout utils/io.py:145-220 reads.  layout "own": every pattern
    has its own graphs (graphs/P_i/G_i_k.gml, split by suffix % 10); "shared": all patterns share graphs/G_k.gml
    (split by suffix % 3).  GML text in python-igraph's writer layout (brackets on their own lines).  Returns the three
    directory paths."""
    from dummynode4graphlearning_b200.subgraph_isomorphism import io as sio

    def one(b, i):
        n0, n1, e0, e1 = int(b["node_ptr"][i]), int(b["node_ptr"][i + 1]), int(b["edge_ptr"][i]), int(b["edge_ptr"][i + 1])
        return dict(num_nodes=n1 - n0, src=b["src"][e0:e1].astype(np.int64) - n0, dst=b["dst"][e0:e1].astype(np.int64) - n0,
                    vid=b["vid"][n0:n1], vlabel=b["vlabel"][n0:n1], elabel=b["elabel"][e0:e1],
                    ekey=np.zeros(e1 - e0, np.int64))

    pd, gd, md = (os.path.join(root, x) for x in ("patterns", "graphs", "metadata"))
    for d in (pd, gd, md):
        os.makedirs(d, exist_ok=True)
    B = int(pattern_b["num_graphs"])
    if layout == "own":
        for i in range(B):
            p = "P_%d" % i
            sio.write_gml_graph(os.path.join(pd, p + ".gml"), one(pattern_b, i), creator)
            os.makedirs(os.path.join(gd, p), exist_ok=True)
            g = "G_%d_%d" % (i, i)
            sio.write_gml_graph(os.path.join(gd, p, g + ".gml"), one(graph_b, i), creator)
            sio.write_metadata_csv(os.path.join(md, p + ".csv"), [(g, counts[i], mats[i])])
    else:
        for i in range(B):
            sio.write_gml_graph(os.path.join(gd, "G_%d.gml" % i), one(graph_b, i), creator)
        for i in range(min(B, 3)):
            p = "P_%d" % i
            sio.write_gml_graph(os.path.join(pd, p + ".gml"), one(pattern_b, i), creator)
            # ground truth is per (pattern, graph): reuse sample i's matrix shape with zero rows where it does not fit
            rows = [("G_%d" % k, counts[k], mats[i] if k == i else np.zeros((0, mats[i].shape[1]), np.int64))
                    for k in range(B)]
            sio.write_metadata_csv(os.path.join(md, p + ".csv"), rows)
    return pd, gd, md


def ref_load_data(pattern_dir, graph_dir, metadata_dir):
    """the reference's utils/io.py:load_data (igraph.read = the shim's GML reader), graphs flattened to dicts of arrays."""
    io = refload.subgraph().io

    def flat(g):
        el = g.get_edgelist()
        return dict(num_nodes=g.vcount(), src=np.asarray([e[0] for e in el], np.int64),
                    dst=np.asarray([e[1] for e in el], np.int64),
                    vid=np.asarray(g.vs["id"], np.int64), vlabel=np.asarray(g.vs["label"], np.int64),
                    elabel=np.asarray(g.es["label"] if el else [], np.int64),
                    ekey=np.asarray(g.es["key"] if el else [], np.int64))

    data, shared = io.load_data(pattern_dir, graph_dir, metadata_dir, num_workers=1)
    out = {}
    for split, xs in data.items():
        out[split] = [dict(id=x["id"], pattern=flat(x["pattern"]), graph=flat(x["graph"]), counts=x["counts"],
                           subisomorphisms=np.asarray(x["subisomorphisms"])) for x in xs]
    return out, shared


class _OneBatchLoader:
    """the only things train_epoch asks of its DataLoader: len(), iteration, and a .dataset that is not an EdgeSeqDataset."""

    def __init__(self, batch):
        self.batch, self.dataset = batch, object()

    def __len__(self):
        return 1

    def __iter__(self):
        return iter([self.batch])


class _SnapshotOptimizer:
    """records the parameter gradients train_epoch has accumulated when it calls step() (it zeroes them right after)."""

    def __init__(self, model):
        self.model, self.grads = model, None

    def step(self):
        self.grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in self.model.named_parameters()}

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None


def ref_train_epoch(model, pattern_g, graph_g, counts, node_weights, edge_weights, config):
    """one mini-batch through the reference's own train_epoch (train.py:596-1000, executed verbatim): returns
    (eval metric, bp_loss, {parameter name: gradient after clipping}).  config: bp_loss, eval_metric, neg_pred_slp,
    match_loss_w, match_reg_w, rep_reg_w, max_grad_norm (numbers, not schedules)."""
    import gc
    import torch.nn as nn
    import torch.nn.functional as F
    sub = refload.subgraph()
    te = refload._extract_functions(os.path.join(refload._SUB, "train.py"), ["train_epoch"],
                                    extra={"np": np, "F": F, "nn": nn, "gc": gc}).train_epoch
    cfg = dict(train_epochs=1, lr=1e-3, train_grad_steps=1, train_log_steps=1, eval_metric="MAE", bp_loss="MSE",
               neg_pred_slp=0.01, match_loss_w=0.0, match_reg_w=0.0, rep_reg_w=0.0, max_grad_norm=0.0)
    cfg.update(config)
    opt = _SnapshotOptimizer(model)
    opt.zero_grad()
    batch = (["x"] * len(counts), pattern_g, graph_g, th.as_tensor(counts), (node_weights, edge_weights))
    metric, loss = te(model, opt, None, "train", _OneBatchLoader(batch), th.device("cpu"), cfg, 0, None, None)
    assert sub is not None
    return metric, loss, opt.grads


def ref_sub_conjugate(b):
    """reference convert_conjugate_graph, DGL branch (utils/graph.py:77-175)."""
    gu = refload.subgraph().graph_utils
    return dgl_list_to_batch([gu.convert_conjugate_graph(g) for g in batch_to_dgl_list(b)])


def ref_process_model_config(config):
    return refload.subgraph().train_funcs.process_model_config(config)


# ------------------------------------------------------------------------------------------
# model-level drivers
def dgl_batched(b):
    """one batched fake DGLGraph with the frames the reference models read:
    ndata id,label[,is_dummy],in_deg,out_deg ; edata id,label[,is_dummy,is_reversed]."""
    C = refload.subgraph().constants
    gs = batch_to_dgl_list(b)
    for g in gs:
        g.ndata[C.INDEGREE] = g.in_degrees()      # calculate_degrees, train.py:477-497 (post-augmentation)
        g.ndata[C.OUTDEGREE] = g.out_degrees()
    return fake_dgl.batch(gs)


def counting_kwargs(model_cfg, **over):
    """constructor kwargs as train.py:1401 passes them (process_model_config'ed maxima + CLI defaults)."""
    kw = dict(model_cfg)
    kw.update(hid_dim=64, rep_num_graph_layers=3, rep_num_pattern_layers=3, rep_act_func="leaky_relu",
              pred_act_func="leaky_relu", pred_net="SumPredictNet", pred_hid_dim=64, emb_net="Equivariant",
              enc_net="Multihot", filter_net="ScalarFilter", pred_with_enc=True, pred_with_deg=True,
              rep_rgin_regularizer="bdd", rep_rgin_num_bases=4, rep_rgin_num_mlp_layers=2,
              rep_dmpnn_num_mlp_layers=2, pred_return_weights="node", rep_residual=True, rep_dropout=0.0,
              pred_dropout=0.0, share_rep_net=True, share_enc_net=True)
    kw.update(over)
    return kw


def ref_counting_model(name, kw, seed=0):
    """the reference's own RGIN / DMPNN class, pred_fc2 / weight_fc2 re-randomised (they are zero-initialised,
    pred.py:50,53, which would make every output 0 -- SURVEY.md App. A-8)."""
    ns = refload.subgraph()
    th.manual_seed(seed)
    model = {"RGIN": ns.rgin.RGIN, "DMPNN": ns.dmpnn.DMPNN, "RGCN": ns.rgcn.RGCN, "CompGCN": ns.compgcn.CompGCN}[name](**kw)
    with th.no_grad():
        for n, p in model.named_parameters():
            if "pred_fc2" in n or "weight_fc2" in n:
                p.normal_(0.0, 0.1)
            if n.endswith("bias") and p.dim() == 1 and "fc" not in n:
                p.normal_(0.0, 0.05)   # rep-layer biases are zero-initialised too; make them count
    return model


def ref_classifier(name, args, seed=0):
    ns = refload.classification()
    th.manual_seed(seed)
    return {"GIN": ns.gconv.GIN, "RGIN": ns.rgconv.RGIN}[name](args)
