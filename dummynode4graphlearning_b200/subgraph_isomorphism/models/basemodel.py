"""Model skeleton of the counting networks: encode -> filter -> embed -> represent -> readout / predict.

Drop-in for the GraphAdj slices of ``subgraph_isomorphism/models/basemodel.py``: ``GraphAdjModel`` (:629-982,
node stream only -- RGIN) and ``GraphAdjModelV2`` (:985-1703, node + edge streams -- DMPNN).  Constructor
kwargs, sub-module names (hence state_dict keys), parameter sharing
(``share_enc_net / share_emb_net / share_rep_net``) and the ``OutputDict`` key set are the reference's.

What changed is HOW the ragged <-> padded plumbing runs: the reference loops over the batch in Python
(``split_and_batchify_graph_feats``, ``batch_convert_len_to_mask``, ``th.cat([gate[i, -len_i:] ...])`` --
utils/dl.py:51-127, basemodel.py:845) with host syncs; here masks, gates and the left-padded readout tensors
are produced by single CUDA kernels on the batched graph (``ops.pad_segments``, ``ops.label_filter_gate``).
"""
from collections import OrderedDict

import os

import torch as th
import torch.nn as nn

from ... import ops
from ..utils import OutputDict, expand_dimensions
from .embed import (EquivariantEmbedding, MultihotEmbedding, NormalEmbedding, OrthogonalEmbedding,
                    PositionEmbedding, UniformEmbedding, get_enc_len)
from .filter import ScalarFilter
from .pred import MaxPredictNet, MeanPredictNet, SumPredictNet

_EMB = {"Orthogonal": OrthogonalEmbedding, "Normal": NormalEmbedding, "Uniform": UniformEmbedding,
        "Equivariant": EquivariantEmbedding}
_PRED = {"MeanPredictNet": MeanPredictNet, "SumPredictNet": SumPredictNet, "MaxPredictNet": MaxPredictNet}


def _padded_mask(g, kind, dummy=True, reversed_=False):
    """(B, Lmax) bool: True on real (non-padded) rows, minus dummy rows (basemodel.py:905-912) and, for
    edges, minus reversed edges (:1562-1571).  One pad kernel on a ones column."""
    frame, ptr, L = (g.ndata, g.node_ptr, g.padded_num_nodes()) if kind == "node" else \
                    (g.edata, g.edge_ptr, g.padded_num_edges())
    n = g.number_of_nodes() if kind == "node" else g.number_of_edges()
    drop = None
    if dummy and "is_dummy" in frame:
        drop = frame["is_dummy"].view(-1).bool()
    if reversed_ and "is_reversed" in frame:
        r = frame["is_reversed"].view(-1).bool()
        drop = r if drop is None else (drop | r)
    ones = g.cached("ones_" + kind, lambda: th.ones((n, 1), dtype=th.float32, device=g.device))
    return ops.pad_segments(ones, ptr, L, drop).view(g.batch_size, L) > 0.5, drop


class _CountingBase(nn.Module):
    """shared constructor plumbing (BaseModel.__init__, basemodel.py:22-59)."""

    # The counting models' wide products (relation table, loop, P|Q, T: section 4 K7 of DESIGN.md) take the library fp32 GEMM
    # unless this is True (DN4GL_COUNTING_GEMM_TC=1 flips the default).  These networks have no normalisation layer and sum
    # over 512-node graphs: the tensor core's truncating (biased) accumulation adds 1e-6 .. 5e-6 to the relative error of the
    # loss, and on an ill-conditioned batch -- where the all-library GPU path is already 8e-6 from float64 -- that is 1.2e-5,
    # over the 1e-5 bar (profiles/r4n_counting_loss_error_sweep.txt).  Speed is a wash either way
    # (profiles/r4k_models_native_gemm_vs_library.txt), so parity decides.
    tensor_core_gemm = os.environ.get("DN4GL_COUNTING_GEMM_TC", "0") == "1"

    def __call__(self, *args, **kwargs):
        with ops.gemm_tensor_cores(self.tensor_core_gemm):
            return super().__call__(*args, **kwargs)

    has_edge_stream = False

    def __init__(self, **kw):
        super().__init__()
        for k in ("max_ngv", "max_ngvl", "max_nge", "max_ngel", "max_npv", "max_npvl", "max_npe", "max_npel"):
            setattr(self, k, kw[k])
        self.base = kw.get("base", 2)
        self.hid_dim = kw.get("hid_dim", 64)
        self.share_emb_net = kw.get("share_emb_net", True)
        self.share_enc_net = kw.get("share_enc_net", True)
        self.share_rep_net = kw.get("share_rep_net", True)
        self.rep_residual = kw.get("rep_residual", True)
        self.pred_with_enc = kw.get("pred_with_enc", False)
        self.pred_with_deg = kw.get("pred_with_deg", False)
        self.add_node_id = kw.get("add_node_id", kw.get("gnn_add_node_id", False))
        self.add_edge_id = kw.get("add_edge_id", kw.get("gnn_add_edge_id", False))
        self.node_pred = kw.get("node_pred", True)
        self.edge_pred = kw.get("edge_pred", True)

        self.g_enc_net = self.create_enc_net(type="graph", **kw)
        self.p_enc_net = self.create_enc_net(type="pattern", **kw)
        self.filter_net = self.create_filter_net(**kw)
        self.g_emb_net = self.create_emb_net(type="graph", **kw)
        self.p_emb_net = self.create_emb_net(type="pattern", **kw)
        self.g_rep_net = self.create_rep_net(type="graph", **kw)
        self.p_rep_net = self.create_rep_net(type="pattern", **kw)
        self.pred_net = self.create_pred_net(**kw)

    # ---- encoders ------------------------------------------------------------------------------
    def _enc_keys(self):
        return ("v", "vl", "el") if self.has_edge_stream else ("v", "vl")

    def _max_of(self, type, key):
        side = "g" if type == "graph" else "p"
        return getattr(self, {"v": "max_n%sv", "vl": "max_n%svl", "el": "max_n%sel"}[key] % side)

    def create_enc_net(self, type, **kw):
        kind = kw.get("enc_net", "Multihot")
        if kind not in ("Multihot", "Position"):
            raise NotImplementedError(kind)            # basemodel.py:649-650
        if type == "pattern" and self.share_enc_net:
            return self.g_enc_net
        if kind == "Multihot":
            enc = OrderedDict((k, MultihotEmbedding(self._max_of(type, k), self.base)) for k in self._enc_keys())
        else:                                          # sinusoid tables of the same width (basemodel.py:642-646)
            enc = OrderedDict((k, PositionEmbedding(get_enc_len(self._max_of(type, k) - 1, self.base) * self.base,
                                                    self._max_of(type, k))) for k in self._enc_keys())
        for net in enc.values():
            net.weight.requires_grad = False
        return nn.ModuleDict(enc)

    def get_graph_enc_dims(self):
        return OrderedDict((k, get_enc_len(self._max_of("graph", k) - 1, self.base) * self.base) for k in self._enc_keys())

    def get_pattern_enc_dims(self):
        if self.share_enc_net:
            return self.get_graph_enc_dims()
        return OrderedDict((k, get_enc_len(self._max_of("pattern", k) - 1, self.base) * self.base)
                           for k in self._enc_keys())

    def create_filter_net(self, **kw):
        f = kw.get("filter_net", "None")
        if f == "None":
            return None
        if f == "ScalarFilter":
            return nn.ModuleDict(OrderedDict((k, ScalarFilter()) for k in self._enc_keys() if k != "v"))
        raise ValueError(f)

    def create_emb_net(self, type, **kw):
        # NB: unlike the encoders the reference never aliases p_emb_net to g_emb_net (share_emb_net is stored but
        # unused for GraphAdj models, basemodel.py:51-52,69-91) -- two independent embedding nets.
        dims = self.get_graph_enc_dims() if type == "graph" else self.get_pattern_enc_dims()
        cls = _EMB[kw.get("emb_net", "Orthogonal")]
        emb = nn.ModuleDict(OrderedDict((k, cls(v, self.hid_dim)) for k, v in dims.items()))
        if self.has_edge_stream:  # V2 rescales for the multi-hot width (basemodel.py:1086-1090)
            with th.no_grad():
                for k in emb:
                    emb[k].weight.div_(dims[k] // self.base)
        return emb

    def _make_pred(self, rep_dim, return_weights, **kw):
        name = kw.get("pred_net", "SumPredictNet")
        if name not in _PRED:
            raise NotImplementedError("%s: attention / memory heads are outside the message-passing hot path" % name)
        return _PRED[name](rep_dim, hidden_dim=kw.get("pred_hid_dim", 64), act_func=kw.get("pred_act_func", "relu"),
                           dropout=kw.get("pred_dropout", 0.0), return_weights=return_weights)

    def refine_node_weights(self, weights, use_max=False):
        return weights

    def refine_edge_weights(self, weights, use_max=False):
        return weights

    # ---- growing a trained model to larger data-set maxima (BaseModel.expand, basemodel.py:167-219) ------------
    def expand(self, **kw):
        """what ``train.py:1399`` calls on every loaded checkpoint: the maxima become max(old, new); encoder tables are
        rebuilt; embedding nets and (with pred_with_enc) the head are rebuilt for the wider encodings and inherit the
        trained values in their trailing corner (``expand_dimensions(pre_pad=True)``), zeros elsewhere.  The
        representation nets are left untouched, as in the reference.  Two reference behaviours kept on purpose: with
        ``share_emb_net`` (default) the pattern embedding net BECOMES the graph's after an expand although the two are
        independent at construction; on failure the maxima are rolled back and the exception re-raised."""
        if "base" in kw and kw["base"] != self.base:
            raise ValueError
        kw = dict(kw)
        names = ("max_npv", "max_npvl", "max_npe", "max_npel", "max_ngv", "max_ngvl", "max_nge", "max_ngel")
        bak = {k: getattr(self, k) for k in names}
        for k in names:
            setattr(self, k, max(kw.get(k, -1), bak[k]))
        try:
            self.g_enc_net = self.create_enc_net(type="graph", **kw)
            self.p_enc_net = self.g_enc_net if self.share_enc_net else self.create_enc_net(type="pattern", **kw)
            new_filter = self.create_filter_net(**kw)
            expand_dimensions(self.filter_net, new_filter, pre_pad=True)   # raises for filter_net "None", as upstream
            self.filter_net = new_filter
            new_g_emb = self.create_emb_net(type="graph", **kw)
            expand_dimensions(self.g_emb_net, new_g_emb, pre_pad=True)
            self.g_emb_net = new_g_emb
            if self.share_emb_net:
                self.p_emb_net = self.g_emb_net
            else:
                new_p_emb = self.create_emb_net(type="pattern", **kw)
                expand_dimensions(self.p_emb_net, new_p_emb, pre_pad=True)
                self.p_emb_net = new_p_emb
            if self.pred_with_enc:
                new_pred = self.create_pred_net(**kw)
                expand_dimensions(self.pred_net, new_pred, pre_pad=True)
                self.pred_net = new_pred
        except Exception:
            for k, v in bak.items():
                setattr(self, k, v)
            raise

    # ---- shared forward pieces ---------------------------------------------------------------------
    def _encode(self, net, g):
        enc = OrderedDict(v=net["v"](g.ndata["id"].view(-1)), vl=net["vl"](g.ndata["label"].view(-1)))
        if self.has_edge_stream:
            enc["el"] = net["el"](g.edata["label"].view(-1))
            if self.add_edge_id:
                u, v = g.all_edges(form="uv", order="eid")
                enc["src"], enc["dst"] = enc["v"][u], enc["v"][v]
        return enc

    def _embed(self, net, enc):
        v_emb = net["vl"](enc["vl"])
        if self.add_node_id:
            v_emb = v_emb + net["v"](enc["v"])
        if not self.has_edge_stream:
            return v_emb
        e_emb = net["el"](enc["el"])
        if self.add_edge_id:
            e_emb = e_emb + net["v"](enc["src"]) + net["v"](enc["dst"])
        return v_emb, e_emb

    def _node_readout(self, g, enc, rep, drop):
        feats = []
        if self.pred_with_enc:
            feats += [enc["v"], enc["vl"]]
        if self.pred_with_deg:
            feats += [g.out_degrees().float().view(-1, 1), g.in_degrees().float().view(-1, 1)]
        return th.cat(feats + [rep], dim=-1) if feats else rep

    def _edge_readout(self, g, enc, rep, drop):
        u, v = g.all_edges(form="uv", order="eid")
        feats = []
        if self.pred_with_enc:
            feats += [enc["v"][u], enc["v"][v], enc["vl"][u], enc["el"], enc["vl"][v]]
        if self.pred_with_deg:
            feats += [g.out_degrees().float().view(-1, 1)[u], g.in_degrees().float().view(-1, 1)[v]]
        return th.cat(feats + [rep], dim=-1) if feats else rep

    @staticmethod
    def _head(net, kind, pattern, p_x, p_drop, p_mask, graph, g_x, g_drop, g_mask):
        """one PredictNet on the flat readout rows: unpadded when the head supports it (pred.py docstring), else through the
        reference's left-padded (B, L, rep) tensors."""
        p_ptr, Lp = (pattern.node_ptr, pattern.padded_num_nodes()) if kind == "node" else (pattern.edge_ptr, pattern.padded_num_edges())
        g_ptr, Lg = (graph.node_ptr, graph.padded_num_nodes()) if kind == "node" else (graph.edge_ptr, graph.padded_num_edges())
        if net.supports_ragged() and p_x.is_cuda:
            return net.forward_ragged(p_x, p_ptr, p_drop, p_mask, Lp, g_x, g_ptr, g_drop, g_mask, Lg)
        return net(ops.pad_segments(p_x, p_ptr, Lp, p_drop), p_mask, ops.pad_segments(g_x, g_ptr, Lg, g_drop), g_mask)


class GraphAdjModel(_CountingBase):
    has_edge_stream = False

    def create_pred_net(self, **kw):
        return self._make_pred(self.get_rep_dim(), "node" in kw.get("pred_return_weights", "none"), **kw)

    def get_graph_enc_dim(self):
        return sum(self.get_graph_enc_dims().values())

    def get_pattern_enc_dim(self):
        return sum(self.get_pattern_enc_dims().values())

    def get_rep_dim(self):
        return self.hid_dim + (self.get_graph_enc_dim() if self.pred_with_enc else 0) + (2 if self.pred_with_deg else 0)

    def get_filter_gate(self, pattern, graph):
        if self.filter_net is None or len(self.filter_net) == 0:
            return None
        return self.filter_net["vl"].gate_from_graphs(pattern, graph, "node")

    def get_pattern_enc(self, pattern):
        return self._encode(self.p_enc_net, pattern)

    def get_graph_enc(self, graph):
        return self._encode(self.g_enc_net, graph)

    def get_pattern_emb(self, p_enc):
        return self._embed(self.p_emb_net, p_enc)

    def get_graph_emb(self, g_enc):
        return self._embed(self.g_emb_net, g_enc)

    def get_subiso_pred(self, p_v_rep, p_v_mask, g_v_rep, g_v_mask):
        """the reference's signature: LEFT-PADDED (B, L, rep) tensors (basemodel.py:958-964)"""
        v_pred_c, v_pred_w = self.pred_net(p_v_rep, p_v_mask, g_v_rep, g_v_mask)
        return v_pred_c, (v_pred_w, None)

    def forward(self, pattern, graph):
        vl_gate = self.get_filter_gate(pattern, graph)
        p_enc = self.get_pattern_enc(pattern)
        p_v_emb = self.get_pattern_emb(p_enc)
        p_v_rep = self.get_pattern_rep(pattern, p_v_emb)
        g_enc = self.get_graph_enc(graph)
        g_v_emb = self.get_graph_emb(g_enc)
        g_v_rep = self.get_graph_rep(graph, g_v_emb, gate=vl_gate)

        p_v_mask, p_drop = _padded_mask(pattern, "node")
        g_v_mask, g_drop = _padded_mask(graph, "node")
        p_v_output = self._node_readout(pattern, p_enc, p_v_rep, p_drop)
        g_v_output = self._node_readout(graph, g_enc, g_v_rep, g_drop)
        pred_c, pred_v = self._head(self.pred_net, "node", pattern, p_v_output, p_drop, p_v_mask, graph, g_v_output, g_drop, g_v_mask)
        pred_e = None
        return OutputDict(
            p_v_emb=p_v_emb, p_e_emb=None, g_v_emb=g_v_emb, g_e_emb=None,
            p_v_rep=p_v_rep, p_e_rep=None, g_v_rep=g_v_rep, g_e_rep=None,
            p_v_mask=p_v_mask, p_e_mask=None, g_v_mask=g_v_mask, g_e_mask=None,
            pred_c=pred_c, pred_v=pred_v, pred_e=pred_e)


class GraphAdjModelV2(_CountingBase):
    has_edge_stream = True

    def create_pred_net(self, **kw):
        rep_v, rep_e = self.get_rep_dim()
        rw = kw.get("pred_return_weights", "none")
        return nn.ModuleDict({
            "v": self._make_pred(rep_v, "node" in rw, **kw) if self.node_pred else None,
            "e": self._make_pred(rep_e, "edge" in rw, **kw) if self.edge_pred else None,
        })

    def get_graph_enc_dim(self):
        d = self.get_graph_enc_dims()
        return d["v"] + d["vl"], (d["v"] + d["vl"]) * 2 + d["el"]

    def get_pattern_enc_dim(self):
        d = self.get_pattern_enc_dims()
        return d["v"] + d["vl"], (d["v"] + d["vl"]) * 2 + d["el"]

    def get_rep_dim(self):
        rv = re = self.hid_dim
        if self.pred_with_enc:
            ev, ee = self.get_graph_enc_dim()
            rv, re = rv + ev, re + ee
        if self.pred_with_deg:
            rv, re = rv + 2, re + 2
        return rv, re

    def get_filter_gate(self, pattern, graph):
        if self.filter_net is None or len(self.filter_net) == 0:
            return None, None
        return (self.filter_net["vl"].gate_from_graphs(pattern, graph, "node"),
                self.filter_net["el"].gate_from_graphs(pattern, graph, "edge"))

    def get_pattern_enc(self, pattern):
        return self._encode(self.p_enc_net, pattern)

    def get_graph_enc(self, graph):
        return self._encode(self.g_enc_net, graph)

    def get_pattern_emb(self, p_enc):
        return self._embed(self.p_emb_net, p_enc)

    def get_graph_emb(self, g_enc):
        return self._embed(self.g_emb_net, g_enc)

    def _mix(self, v_c, e_c, g_v_mask, g_e_mask):
        if self.node_pred and self.edge_pred:   # length-weighted mix (basemodel.py:1506-1512)
            g_v_len = g_v_mask.float().sum(dim=1).view(-1, 1)
            g_e_len = g_e_mask.float().sum(dim=1).view(-1, 1)
            g_len = g_v_len + g_e_len
            return (g_v_len / g_len) * v_c + (g_e_len / g_len) * e_c
        if self.node_pred:
            return v_c
        if self.edge_pred:
            return e_c
        raise ValueError

    def get_subiso_pred(self, p_v_rep, p_v_mask, p_e_rep, p_e_mask, g_v_rep, g_v_mask, g_e_rep, g_e_mask):
        """the reference's signature: LEFT-PADDED (B, L, rep) tensors (basemodel.py:1497-1518)"""
        v_c = v_w = e_c = e_w = None
        if self.node_pred:
            v_c, v_w = self.pred_net["v"](p_v_rep, p_v_mask, g_v_rep, g_v_mask)
        if self.edge_pred:
            e_c, e_w = self.pred_net["e"](p_e_rep, p_e_mask, g_e_rep, g_e_mask)
        return self._mix(v_c, e_c, g_v_mask, g_e_mask), (v_w, e_w)

    def forward(self, pattern, graph):
        vl_gate, el_gate = self.get_filter_gate(pattern, graph)
        p_enc = self.get_pattern_enc(pattern)
        p_v_emb, p_e_emb = self.get_pattern_emb(p_enc)
        p_v_rep, p_e_rep = self.get_pattern_rep(pattern, p_v_emb, p_e_emb)
        g_enc = self.get_graph_enc(graph)
        g_v_emb, g_e_emb = self.get_graph_emb(g_enc)
        g_v_rep, g_e_rep = self.get_graph_rep(graph, g_v_emb, g_e_emb, v_gate=vl_gate, e_gate=el_gate)

        p_v_mask, p_v_drop = _padded_mask(pattern, "node")
        g_v_mask, g_v_drop = _padded_mask(graph, "node")
        p_e_mask, p_e_drop = _padded_mask(pattern, "edge", reversed_=True)
        g_e_mask, g_e_drop = _padded_mask(graph, "edge", reversed_=True)

        v_c = pred_v = e_c = pred_e = None
        if self.node_pred:
            v_c, pred_v = self._head(self.pred_net["v"], "node", pattern, self._node_readout(pattern, p_enc, p_v_rep, p_v_drop),
                                     p_v_drop, p_v_mask, graph, self._node_readout(graph, g_enc, g_v_rep, g_v_drop), g_v_drop, g_v_mask)
        if self.edge_pred:
            e_c, pred_e = self._head(self.pred_net["e"], "edge", pattern, self._edge_readout(pattern, p_enc, p_e_rep, p_e_drop),
                                     p_e_drop, p_e_mask, graph, self._edge_readout(graph, g_enc, g_e_rep, g_e_drop), g_e_drop, g_e_mask)
        pred_c = self._mix(v_c, e_c, g_v_mask, g_e_mask)
        return OutputDict(
            p_v_emb=p_v_emb, p_e_emb=p_e_emb, g_v_emb=g_v_emb, g_e_emb=g_e_emb,
            p_v_rep=p_v_rep, p_e_rep=p_e_rep, g_v_rep=g_v_rep, g_e_rep=g_e_rep,
            p_v_mask=p_v_mask, p_e_mask=p_e_mask, g_v_mask=g_v_mask, g_e_mask=g_e_mask,
            pred_c=pred_c, pred_v=pred_v, pred_e=pred_e)
