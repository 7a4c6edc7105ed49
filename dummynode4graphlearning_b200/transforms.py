"""GPU graph transforms (dummy-node augmentation, edge-to-vertex transform, PyG canonicalisation).

Host-side mirror of the reference functions, operating on whole mini-batches in HBM:

=========================================  ==================================================
reference (file:line)                      here
=========================================  ==================================================
load_graph_data_from_TUDatadir(with_dummy) ``tu_add_dummy``      (tu_data_processing.py:186-214)
convert_conjugate_graph_forward            ``tu_conjugate``      (tu_data_processing.py:223-338)
PyG read_tu_data + set_dummy_flags         ``pyg_canonicalize``  (graph_neural_networks/dataset.py:118-151)
add_dummy_nodes_edges (GraphAdj branch)    ``sub_add_dummy``     (subgraph_isomorphism/train.py:404-474)
convert_conjugate_graph / _to_conjugate    ``sub_conjugate``     (utils/graph.py:77-175, train.py:564-593)
process_model_config                       ``process_model_config`` (train.py:38-81)
=========================================  ==================================================

Batches are dicts of int32 device tensors in the flat layout of ``synth.py``:
``node_ptr, edge_ptr, src, dst`` (global ids) + attribute columns.  All index outputs are
bit-exact with the reference (tests/test_transforms_gpu.py).
"""
import math
from copy import deepcopy

import numpy as np
import torch

from ._lib import lib, ptr
from .graph import _stream, build_csr, error_flag, require_cuda


class LazyDict(dict):
    """batch dict whose derived columns (ids, flags, int64 views, ...) are computed on first access: the train step of
    GIN reads only a few of the columns the reference's files carry, and every skipped column is a few launches."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._lazy = {}

    def lazy(self, key, fn):
        self._lazy[key] = fn

    def __missing__(self, key):
        fn = self._lazy.pop(key, None)
        if fn is None:
            raise KeyError(key)
        v = fn()
        self[key] = v
        return v

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def get(self, key, default=None):
        return self[key] if key in self else default

    def pop(self, key, *default):
        self._lazy.pop(key, None)
        return dict.pop(self, key, *default)

    def materialize(self):
        for k in list(self._lazy):
            self[k]
        return self

    def keys(self):
        self.materialize()
        return dict.keys(self)

    def items(self):
        self.materialize()
        return dict.items(self)

    def values(self):
        self.materialize()
        return dict.values(self)

    def __iter__(self):
        self.materialize()
        return dict.__iter__(self)


def _take(t, idx_i32):
    """t[idx] for an int32 index tensor (one launch, no int64 copy of the index)."""
    return torch.index_select(t, 0, idx_i32)


def to_device(b, device):
    """numpy/tensor batch dict -> int32 (float32 for *attr) device tensors."""
    dev = torch.device(device)
    out = {}
    for k, v in b.items():
        if k in ("num_graphs", "has_edge_labels", "max_graph_nodes", "conj_sizes"):
            out[k] = v
        elif k.endswith("attr"):
            out[k] = torch.as_tensor(v).to(dev, torch.float32).contiguous()
        elif k == "y":
            out[k] = torch.as_tensor(v).to(dev, torch.int64)
        else:
            out[k] = torch.as_tensor(v).to(dev, torch.int32).contiguous()
    return out


def _empty_i32(n, dev):
    return torch.empty(int(n), dtype=torch.int32, device=dev)


def _local_ids(ptrs, total, dev):
    """position within the segment for every element (ID = range(...), tu_data_processing.py:213-214)."""
    counts = (ptrs[1:] - ptrs[:-1]).long()
    starts = torch.repeat_interleave(ptrs[:-1].long(), counts, output_size=int(total))
    return (torch.arange(int(total), device=dev) - starts).to(torch.int32)


def tu_add_dummy(b):
    """+1 dummy node (LABEL 0) and 2n interleaved dummy edges per graph."""
    require_cuda(b["src"], "batch")
    dev = b["src"].device
    B, N, E = int(b["num_graphs"]), int(b["vlabel"].numel()), int(b["src"].numel())
    o = LazyDict(num_graphs=B, node_ptr=_empty_i32(B + 1, dev), edge_ptr=_empty_i32(B + 1, dev),
                 src=_empty_i32(E + 2 * N, dev), dst=_empty_i32(E + 2 * N, dev),
                 vlabel=_empty_i32(N + B, dev), v_is_dummy=_empty_i32(N + B, dev),
                 elabel=_empty_i32(E + 2 * N, dev), e_is_dummy=_empty_i32(E + 2 * N, dev))
    lib().call("dn4gl_tu_add_dummy", B, ptr(b["node_ptr"]), ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["dst"]),
               ptr(b["vlabel"]), ptr(b["elabel"]), N, E,
               ptr(o["node_ptr"]), ptr(o["edge_ptr"]), ptr(o["src"]), ptr(o["dst"]),
               ptr(o["vlabel"]), ptr(o["v_is_dummy"]), ptr(o["elabel"]), ptr(o["e_is_dummy"]), _stream())
    # NB: the thunks must not capture `o` itself (o -> thunk -> o is a reference cycle: the batch's device tensors would
    # live until the cyclic GC runs and the caching allocator would cudaMalloc fresh blocks meanwhile)
    o_np, o_ep = o["node_ptr"], o["edge_ptr"]
    o.lazy("vid", lambda: _local_ids(o_np, N + B, dev))
    o.lazy("eid", lambda: _local_ids(o_ep, E + 2 * N, dev))
    if "vattr" in b:  # dummy node gets attribute 0 (line 191)
        va = torch.zeros(N + B, dtype=torch.float32, device=dev)
        va[o["v_is_dummy"] == 0] = b["vattr"]
        o["vattr"] = va
    if "y" in b:
        o["y"] = b["y"]
    if b.get("max_graph_nodes") is not None:
        o["max_graph_nodes"] = int(b["max_graph_nodes"]) + 1
    if b.get("conj_sizes") is not None:      # host-side size hint for the tu_conjugate that follows (tu_conjugate_sizes)
        o["conj_sizes"] = b["conj_sizes"]
    return o


def tu_conjugate_sizes(raw, with_dummy):
    """Sizes of the edge-to-vertex transform of a RAW TU-flavoured host batch (numpy arrays), computed on the host without
    building it: ``(V', E', nodes of the largest output graph)`` of ``tu_conjugate(tu_add_dummy(raw))`` (with_dummy) or
    ``tu_conjugate(raw)`` (LINE_).  V' = m [+ 1 per graph], E' = sum_v in(v) out(v) [+ 2 m] (tu_data_processing.py:223-338;
    multi-edges count with their multiplicity, a self loop is both an in- and an out-edge of its node).  Put the result
    into the batch as ``conj_sizes`` (a loader knows its batch on the host anyway): ``tu_conjugate`` then allocates its
    outputs from it instead of reading the sizes back from the device, which removes the transform's only host
    synchronisation.  The hint MUST come from this function applied to the same batch: the fill kernel writes E' edges."""
    node_ptr, edge_ptr = np.asarray(raw["node_ptr"], np.int64), np.asarray(raw["edge_ptr"], np.int64)
    src, dst = np.asarray(raw["src"], np.int64), np.asarray(raw["dst"], np.int64)
    B, N, m = len(node_ptr) - 1, int(node_ptr[-1]), int(edge_ptr[-1])
    ind, outd = np.bincount(dst, minlength=N), np.bincount(src, minlength=N)
    pairs = int((ind * outd).sum())
    per_graph = np.diff(edge_ptr)
    if with_dummy:
        has_nodes = np.diff(node_ptr) > 0          # a graph without nodes gets no dummy edges, hence no dummy vertex
        return m + int(has_nodes.sum()), pairs + 2 * m, int((per_graph + has_nodes).max()) if B else 0
    return m, pairs, int(per_graph.max()) if B else 0


def exclusive_scan_(lens_plus_one):
    """in place: x[i] <- sum_{j<i} x[j] over the first n = len - 1 entries, x[n] <- total (dn4gl_exclusive_scan_i32)."""
    L = lib()
    n = int(lens_plus_one.numel()) - 1
    wsb = L.size("dn4gl_scan_workspace_bytes", n)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=lens_plus_one.device)
    L.call("dn4gl_exclusive_scan_i32", ptr(lens_plus_one), ptr(lens_plus_one), n, ptr(ws), wsb, _stream())
    return lens_plus_one


def tu_conjugate_sizes_ex(raw, with_dummy=True):
    """``tu_conjugate_sizes`` plus what the closed-form CONJ_ builder (``tu_conj_structure``) needs to know on the host:
    (V', E', max nodes, every graph has a node, edges are sorted by source)."""
    v, e, mx = tu_conjugate_sizes(raw, with_dummy)
    node_ptr, src = np.asarray(raw["node_ptr"], np.int64), np.asarray(raw["src"], np.int64)
    return (v, e, mx, bool((np.diff(node_ptr) > 0).all()), bool((np.diff(src) >= 0).all()) if len(src) > 1 else True)


def tu_conj_structure(b, num_node_labels, node_label_min=0):
    """RAW TU-shaped device batch -> the canonical CONJ_ mini-batch (what ``pyg_canonicalize(tu_conjugate(tu_add_dummy(b)))``
    yields for a model that reads x, y and the graph structure: GIN), written in closed form by
    dn4gl_tu_conj_direct_lens / _fill from the raw graphs' two CSRs: no dummy-augmented graph, no candidate list, no
    sort, no compaction (csrc/transforms.cu; tu_data_processing.py:125-338 + dataset.py:151).  Needs the loader's hint
    ``b["conj_sizes"]`` from ``tu_conjugate_sizes_ex`` with "every graph has a node" true.  -> graph_classification.data.Batch"""
    from .graph import CSR, HEAVY_THRESHOLD
    from .graph_classification.data import Batch, GraphStructure
    require_cuda(b["src"], "batch")
    L = lib()
    dev = b["src"].device
    B, N, E = int(b["num_graphs"]), int(b["vlabel"].numel()), int(b["src"].numel())
    V2, E2cap, max_nodes, all_nodes, sorted_src = b["conj_sizes"]
    if not all_nodes or V2 != E + B:
        raise ValueError("tu_conj_structure needs graphs with at least one node each")
    out = build_csr(b["src"], b["dst"], N, heavy_threshold=0, sorted_keys=bool(sorted_src))
    inn = build_csr(b["dst"], b["src"], N, heavy_threshold=0)
    V = E + B
    rp_out, rp_in, node_ptr = _empty_i32(V + 1, dev), _empty_i32(V + 1, dev), _empty_i32(B + 1, dev)
    L.call("dn4gl_tu_conj_direct_lens", B, ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["dst"]), ptr(out.row_ptr), ptr(inn.row_ptr),
           E, ptr(rp_out), ptr(rp_in), ptr(node_ptr), _stream())
    exclusive_scan_(rp_out)
    exclusive_scan_(rp_in)
    cap = max(int(E2cap), 1)      # E' counts the (e, e) pairs of self-loop edges, which the rows leave out: an upper bound
    col_out, col_in, vlabel = _empty_i32(cap, dev), _empty_i32(cap, dev), _empty_i32(V, dev)
    el = b.get("elabel") if b.get("has_edge_labels", True) else None
    e2g = _empty_i32(max(E, 1), dev)
    if E > 0:
        L.call("dn4gl_segment_ids_i32", ptr(b["edge_ptr"]), B, E, ptr(e2g), _stream())
    L.call("dn4gl_tu_conj_direct_fill", B, ptr(b["edge_ptr"]), ptr(e2g), ptr(b["src"]), ptr(b["dst"]), ptr(el), ptr(out.row_ptr), ptr(out.eid),
           ptr(inn.row_ptr), ptr(inn.eid), E, ptr(rp_out), ptr(rp_in), ptr(col_out), ptr(col_in), ptr(vlabel), _stream())
    csrs = []
    for rp, col in ((rp_in, col_in), (rp_out, col_out)):
        c = CSR(rp, col, None, V, cap)
        hcap = cap // HEAVY_THRESHOLD + 1
        c.heavy_rows, c.heavy_count, c.heavy_thr = _empty_i32(hcap, dev), _empty_i32(1, dev), HEAVY_THRESHOLD
        L.call("dn4gl_collect_heavy_rows", ptr(rp), V, HEAVY_THRESHOLD, ptr(c.heavy_rows), hcap, ptr(c.heavy_count), _stream())
        c.seg_ptr, c.max_seg, c.block_diagonal = node_ptr, int(max_nodes), True
        csrs.append(c)
    x = (vlabel.view(-1, 1) == torch.arange(int(node_label_min), int(node_label_min) + int(num_node_labels),
                                            dtype=vlabel.dtype, device=dev)).to(torch.float32)
    def edge_index():      # only a consumer outside the GIN train step asks for it: one device->host read of the edge count
        nnz = int(rp_out[-1].item())
        rows = torch.repeat_interleave(torch.arange(V, device=dev), (rp_out[1:] - rp_out[:-1]).long(), output_size=nnz)
        return torch.stack([rows, col_out[:nnz].long()])

    batch = lambda: torch.repeat_interleave(torch.arange(B, device=dev), (node_ptr[1:] - node_ptr[:-1]).long(), output_size=V)
    data = Batch(x, edge_index, batch, y=b.get("y"), is_dummy_node=lambda: vlabel == 0, node_ptr=node_ptr,
                 max_graph_nodes=int(max_nodes))
    data._structure = GraphStructure.from_csr(csrs[0], csrs[1], V, node_ptr, int(max_nodes))
    return data


def tu_conjugate(b):
    """edge-to-vertex transform of a TU-flavoured batch (raw -> LINE_, dummy-augmented -> CONJ_)."""
    require_cuda(b["src"], "batch")
    L = lib()
    dev = b["src"].device
    B, N, E = int(b["num_graphs"]), int(b["vlabel"].numel()), int(b["src"].numel())
    isd = b.get("e_is_dummy")
    csr_in = build_csr(b["dst"], b["src"], N, heavy_threshold=0)
    ws_bytes = L.size("dn4gl_conj_workspace_bytes", B, N, E)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cand_off, newid = _empty_i32(E + 1, dev), _empty_i32(max(E, 1), dev)
    o_node_ptr, o_edge_ptr = _empty_i32(B + 1, dev), _empty_i32(B + 1, dev)
    L.call("dn4gl_tu_conjugate_count", B, ptr(b["node_ptr"]), ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["dst"]),
           ptr(isd), N, E, ptr(csr_in.row_ptr), ptr(csr_in.eid), ptr(cand_off), ptr(newid),
           ptr(o_node_ptr), ptr(o_edge_ptr), ptr(ws), ws_bytes, _stream())
    hint = b.get("conj_sizes")
    if hint is not None:                      # sizes known on the host (tu_conjugate_sizes): no read-back, no sync
        V2, E2, max_nodes = (int(v) for v in tuple(hint)[:3])      # tu_conjugate_sizes or the first fields of ..._sizes_ex
        flag = error_flag(dev)                # a hint that does not match the device's counts raises the async flag
        wrong = (o_node_ptr[-1:] != V2) | (o_edge_ptr[-1:] != E2)
        flag.copy_(torch.where(wrong, torch.full_like(flag, -5), flag))
    else:
        max_nodes = (o_node_ptr[1:] - o_node_ptr[:-1]).max() if B > 0 else o_node_ptr[-1]
        sizes = torch.stack([o_node_ptr[-1], o_edge_ptr[-1], max_nodes]).cpu()  # the one D2H sync: output sizes
        V2, E2, max_nodes = int(sizes[0]), int(sizes[1]), int(sizes[2])
    o = LazyDict(num_graphs=B, max_graph_nodes=max_nodes, node_ptr=o_node_ptr, edge_ptr=o_edge_ptr,
                 src=_empty_i32(E2, dev), dst=_empty_i32(E2, dev),
                 v_origin=_empty_i32(V2, dev), e_shared=_empty_i32(E2, dev))
    L.call("dn4gl_tu_conjugate_fill", B, ptr(b["node_ptr"]), ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["dst"]),
           ptr(isd), N, E, ptr(csr_in.row_ptr), ptr(csr_in.eid), ptr(cand_off), ptr(newid),
           ptr(o_node_ptr), ptr(o_edge_ptr), ptr(o["src"]), ptr(o["dst"]), ptr(o["v_origin"]), ptr(o["e_shared"]),
           V2, E2, ptr(error_flag(dev)), ptr(ws), ws_bytes, _stream())
    vo, es = o["v_origin"], o["e_shared"]
    # vertex attributes <- original edge attributes (lines 238-242), edge attributes <- shared vertex (322-326);
    # derived lazily: a consumer that only needs the structure and the vertex labels pays for nothing else
    o.lazy("vlabel", lambda: _take(b["elabel"], vo))
    o.lazy("elabel", lambda: _take(b["vlabel"], es))
    o.lazy("vid", lambda: _take(b["eid"] if "eid" in b else _local_ids(b["edge_ptr"], E, dev), vo))
    o.lazy("eid", lambda: _take(b["vid"] if "vid" in b else _local_ids(b["node_ptr"], N, dev), es))
    if isd is not None:
        o.lazy("v_is_dummy", lambda: _take(isd, vo))
        o.lazy("e_is_dummy", lambda: _take(b["v_is_dummy"], es))
    if "vattr" in b:
        o.lazy("eattr", lambda: _take(b["vattr"], es))
    if "y" in b:
        o["y"] = b["y"]
    return o


def sub_add_dummy(b, max_nv, max_nvl, max_ne, max_nel):
    """subgraph-isomorphism dummy augmentation; max_* are the PRE-augmentation maxima passed at
    train.py:1322-1334."""
    require_cuda(b["src"], "batch")
    dev = b["src"].device
    B, N, E = int(b["num_graphs"]), int(b["vlabel"].numel()), int(b["src"].numel())
    o = dict(num_graphs=B, node_ptr=_empty_i32(B + 1, dev), edge_ptr=_empty_i32(B + 1, dev),
             src=_empty_i32(E + 2 * N, dev), dst=_empty_i32(E + 2 * N, dev))
    for k in ("vid", "vlabel", "v_is_dummy"):
        o[k] = _empty_i32(N + B, dev)
    for k in ("eid", "elabel", "e_is_dummy", "e_is_reversed"):
        o[k] = _empty_i32(E + 2 * N, dev)
    lib().call("dn4gl_sub_add_dummy", B, ptr(b["node_ptr"]), ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["dst"]),
               ptr(b["vid"]), ptr(b["vlabel"]), ptr(b["eid"]), ptr(b["elabel"]), ptr(b.get("e_is_reversed")),
               N, E, int(max_nv), int(max_nvl), int(max_ne), int(max_nel),
               ptr(o["node_ptr"]), ptr(o["edge_ptr"]), ptr(o["src"]), ptr(o["dst"]),
               ptr(o["vid"]), ptr(o["vlabel"]), ptr(o["v_is_dummy"]),
               ptr(o["eid"]), ptr(o["elabel"]), ptr(o["e_is_dummy"]), ptr(o["e_is_reversed"]), _stream())
    return o


def sub_add_reversed(b, max_ne, max_nel):
    """add_reversed_edges, GraphAdj branch (train.py:291-345): every graph's edges followed by their reversals with
    id = max_ne + position, label + max_nel, is_reversed = 1.  A batch that already carries ``e_is_reversed`` is returned
    unchanged (train.py:321,334).  max_* are the PRE-augmentation maxima passed at train.py:1315."""
    if "e_is_reversed" in b:
        return b
    require_cuda(b["src"], "batch")
    dev = b["src"].device
    B, E = int(b["num_graphs"]), int(b["src"].numel())
    o = dict(b)
    o["edge_ptr"] = _empty_i32(B + 1, dev)
    for k in ("src", "dst", "eid", "elabel", "e_is_reversed"):
        o[k] = _empty_i32(2 * E, dev)
    lib().call("dn4gl_sub_add_reversed", B, ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["dst"]), ptr(b["eid"]),
               ptr(b["elabel"]), E, int(max_ne), int(max_nel), ptr(o["edge_ptr"]), ptr(o["src"]), ptr(o["dst"]),
               ptr(o["eid"]), ptr(o["elabel"]), ptr(o["e_is_reversed"]), _stream())
    return o


_EDGE_COLUMNS = ("src", "dst", "eid", "elabel", "e_is_dummy", "e_is_reversed", "eattr")


def sub_remove_loops(b):
    """remove_loops, GraphAdj branch (train.py:270-288): edges with u == v dropped, survivors keep their order and
    attribute rows (DGL ``remove_edges``).  One device->host read-back (the surviving edge count)."""
    require_cuda(b["src"], "batch")
    L = lib()
    dev = b["src"].device
    B, E = int(b["num_graphs"]), int(b["src"].numel())
    keep_scan, surv, o_edge_ptr = _empty_i32(E + 1, dev), _empty_i32(max(E, 1), dev), _empty_i32(B + 1, dev)
    wsb = L.size("dn4gl_scan_workspace_bytes", E + 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    L.call("dn4gl_remove_loops_mark", B, ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["dst"]), E, ptr(keep_scan),
           ptr(o_edge_ptr), ptr(surv), ptr(ws), wsb, _stream())
    E2 = int(keep_scan[-1].item())
    surv = surv[:E2]
    o = dict(b)
    o["edge_ptr"] = o_edge_ptr
    for k in _EDGE_COLUMNS:
        if k in b:
            o[k] = _take(b[k], surv)
    return o


def compute_norm(graph, self_loop=True):
    """compute_norm (utils/graph.py:11-38) on a BatchedGraph -> (node_norm (N, 1), edge_norm (E, 1))."""
    in_deg = graph.in_degrees().float()
    if self_loop:
        node_norm = (in_deg + 1).reciprocal().unsqueeze(-1)
    else:
        node_norm = in_deg.reciprocal().masked_fill_(in_deg == 0, 1.0).unsqueeze(-1)
    return node_norm, torch.index_select(node_norm, 0, graph.dst)


def compute_largest_eigenvalues(graph):
    """compute_largest_eigenvalues (utils/graph.py:41-71) for every graph of a BatchedGraph -> (node_eigenv (B,),
    edge_eigenv (B,)) as floats."""
    B = graph.batch_size
    dev = graph.src.device
    ind = graph.cached("in_deg_i32", lambda: graph.in_degrees().to(torch.int32).contiguous())
    outd = graph.cached("out_deg_i32", lambda: graph.out_degrees().to(torch.int32).contiguous())
    ne, ee = _empty_i32(B, dev), _empty_i32(B, dev)
    lib().call("dn4gl_sub_eigen_bounds", B, ptr(graph.edge_ptr), ptr(graph.src), ptr(graph.dst), ptr(ind), ptr(outd),
               ptr(ne), ptr(ee), _stream())
    return ne.float(), ee.float()


def calculate_norms(graph, self_loop=True):
    """calculate_norms (train.py:500-512): ``ndata['norm']`` / ``edata['norm']``."""
    if "norm" not in graph.ndata or "norm" not in graph.edata:
        graph.ndata["norm"], graph.edata["norm"] = compute_norm(graph, self_loop)
    return graph


def calculate_eigenvalues(graph):
    """calculate_eigenvalues (train.py:515-527): per-graph bounds clamped at 1 and repeated over the graph's nodes /
    edges as ``ndata['node_eigenv']`` (N, 1) / ``edata['edge_eigenv']`` (E, 1)."""
    if "node_eigenv" not in graph.ndata or "edge_eigenv" not in graph.edata:
        ne, ee = compute_largest_eigenvalues(graph)
        graph.ndata["node_eigenv"] = torch.repeat_interleave(ne.clamp_min(1.0), graph.batch_num_nodes(),
                                                             output_size=graph.number_of_nodes()).unsqueeze(-1)
        graph.edata["edge_eigenv"] = torch.repeat_interleave(ee.clamp_min(1.0), graph.batch_num_edges(),
                                                             output_size=graph.number_of_edges()).unsqueeze(-1)
    return graph


def sub_conjugate(b, id_bound=None):
    """edge-to-vertex transform of a subgraph-isomorphism batch (``convert_conjugate_graph``, utils/graph.py:77-175,
    as applied by ``convert_to_conjugate``, train.py:564-593): edges with equal ``eid`` merge into one vertex, duplicate
    ``(eid[e'], vlabel[shared], eid[e])`` conjugate edges are dropped keeping the first, and the node / edge attribute
    names are swapped (:155-165).  ``id_bound``: exclusive upper bound of the edge ids (``max_nge`` after augmentation,
    i.e. ``process_model_config(...)["max_nge"]``); read from the data (one extra sync) when omitted."""
    require_cuda(b["src"], "batch")
    L = lib()
    dev = b["src"].device
    B, N, E = int(b["num_graphs"]), int(b["vlabel"].numel()), int(b["src"].numel())
    if id_bound is None:
        id_bound = int(b["eid"].max().item()) + 1 if E else 1
    id_bound = max(int(id_bound), 1)
    csr_in = build_csr(b["dst"], b["src"], N, heavy_threshold=0)
    ws_bytes = L.size("dn4gl_sub_conj_workspace_bytes", B, E, id_bound)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ev, cand_off = _empty_i32(max(E, 1), dev), _empty_i32(E + 1, dev)
    o_node_ptr, o_edge_ptr = _empty_i32(B + 1, dev), _empty_i32(B + 1, dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("dn4gl_sub_conj_count", B, ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["eid"]), N, E, id_bound,
           ptr(csr_in.row_ptr), ptr(ev), ptr(cand_off), ptr(o_node_ptr), ptr(ws), ws_bytes, ptr(err), _stream())
    sizes = torch.stack([o_node_ptr[-1], cand_off[-1], err[0]]).cpu()           # sync 1: V', ncand
    V2, ncand = int(sizes[0]), int(sizes[1])
    if int(sizes[2]) != 0:
        raise RuntimeError("sub_conjugate: an edge id lies outside [0, id_bound=%d)" % id_bound)
    slots = 2
    while slots < 2 * ncand:
        slots *= 2
    table = torch.empty(slots, dtype=torch.int64, device=dev)
    keep_scan = _empty_i32(ncand + 1, dev)
    sws_bytes = L.size("dn4gl_scan_workspace_bytes", ncand + 1)
    sws = torch.empty(sws_bytes, dtype=torch.uint8, device=dev)
    L.call("dn4gl_sub_conj_mark", B, ptr(b["edge_ptr"]), ptr(b["src"]), ptr(b["vlabel"]), E, ptr(csr_in.row_ptr),
           ptr(csr_in.eid), ptr(ev), ptr(cand_off), ncand, ptr(table), slots, ptr(keep_scan), ptr(o_edge_ptr),
           ptr(sws), sws_bytes, _stream())
    E2 = int(o_edge_ptr[-1].item())                                              # sync 2: E'
    o = dict(num_graphs=B, node_ptr=o_node_ptr, edge_ptr=o_edge_ptr, src=_empty_i32(E2, dev), dst=_empty_i32(E2, dev),
             v_origin=_empty_i32(V2, dev), e_shared=_empty_i32(E2, dev))
    L.call("dn4gl_sub_conj_fill", B, ptr(b["src"]), E, id_bound, ptr(csr_in.row_ptr), ptr(csr_in.eid), ptr(ev),
           ptr(cand_off), ptr(keep_scan), ptr(o["src"]), ptr(o["dst"]), ptr(o["v_origin"]), ptr(o["e_shared"]),
           ptr(ws), ws_bytes, _stream())
    vo, es = o["v_origin"].long(), o["e_shared"].long()
    for ek, vk in (("eid", "vid"), ("elabel", "vlabel"), ("e_is_dummy", "v_is_dummy"), ("e_is_reversed", "v_is_reversed")):
        if ek in b:
            o[vk] = b[ek][vo]
    for vk, ek in (("vid", "eid"), ("vlabel", "elabel"), ("v_is_dummy", "e_is_dummy")):
        if vk in b:
            o[ek] = b[vk][es]
    return o


def pyg_canonicalize(b, num_node_labels=None, num_edge_labels=None, node_label_min=None, with_edge_attr=True,
                     defer_count=False):
    """What PyG ``read_tu_data`` + ``PYGDataset.set_dummy_flags`` turn the saved TU files into
    (graph_neural_networks/dataset.py:118-151): one-hot ``x`` (attribute column first, then labels
    shifted to start at 0), ``edge_index`` int64 with self loops removed and coalesced = sorted by
    (row, col) with duplicates merged, ``edge_attr`` (one-hot labels, summed over merged
    duplicates), ``batch``, ``is_dummy_node`` / ``is_dummy_edge``."""
    require_cuda(b["src"], "batch")
    L = lib()
    dev = b["src"].device
    B, N, E = int(b["num_graphs"]), int(b["vlabel"].numel()), int(b["src"].numel())
    csr_out = build_csr(b["src"], b["dst"], N, heavy_threshold=0)
    keep_scan = _empty_i32(E + 1, dev)
    o_src, o_dst, o_first = _empty_i32(E, dev), _empty_i32(E, dev), _empty_i32(E, dev)
    ws_bytes = L.size("dn4gl_coalesce_workspace_bytes", N, E)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    L.call("dn4gl_coalesce", ptr(b["src"]), ptr(b["dst"]), N, E, ptr(csr_out.row_ptr), ptr(csr_out.eid), ptr(keep_scan),
           ptr(o_src), ptr(o_dst), ptr(o_first), ptr(ws), ws_bytes, ptr(error_flag(dev)), _stream())
    defer = bool(defer_count) and not (with_edge_attr and b.get("has_edge_labels", True)) and E > 0
    if defer:
        # the surviving edge count stays on the device: the arrays keep their capacity E, the tail [E'', E) points at a
        # trash row N (dn4gl_pad_tail_i32), and the CSRs are built over N + 1 rows -- no second device->host read-back on
        # the GIN path.  Whoever needs the exact lists (edge_index, first_edge) pays for the read-back lazily.
        L.call("dn4gl_pad_tail_i32", ptr(keep_scan[E:]), E, ptr(o_src), ptr(o_dst), N, _stream())
        count = {}

        def exact():
            if "E2" not in count:
                count["E2"] = int(keep_scan[-1].item())
            return count["E2"]
        E2 = None
    else:
        E2 = int(keep_scan[-1].item())
        o_src, o_dst, o_first = o_src[:E2], o_dst[:E2], o_first[:E2]
    # node features: [attr?, one_hot(label - min)]
    vl = b["vlabel"]
    if node_label_min is not None and num_node_labels is not None:
        vmin = int(node_label_min)      # caller knows the label range: no device->host sync
    else:
        vmin = int(vl.min().item()) if N else 0
    nvl = int(num_node_labels) if num_node_labels is not None else (int(vl.max().item()) - vmin + 1 if N else 0)
    x = (vl.view(-1, 1) == torch.arange(vmin, vmin + nvl, dtype=vl.dtype, device=dev)).to(torch.float32)
    n_attr = 0
    if "vattr" in b:
        x = torch.cat([b["vattr"].view(N, -1), x], dim=1)
        n_attr = x.size(1) - nvl
    out = LazyDict(num_graphs=B, x=x, node_ptr=b["node_ptr"], src=o_src, dst=o_dst,
                   sorted_by_src=True)   # survivors are compacted in (src, dst) order: the by-src CSR needs no sort
    if defer:
        out["trash_row"] = True          # src / dst hold E entries, the last E - E'' of them (N, N)
        out.lazy("first_edge", lambda: o_first[:exact()])
        out.lazy("edge_index", lambda: torch.stack([o_src[:exact()].long(), o_dst[:exact()].long()]))
        out.lazy("num_edges", exact)
    else:
        out["first_edge"] = o_first
        out.lazy("edge_index", lambda: torch.stack([o_src.long(), o_dst.long()]))
    out.lazy("batch", lambda: torch.repeat_interleave(torch.arange(B, device=dev),
                                                      (b["node_ptr"][1:] - b["node_ptr"][:-1]).long(), output_size=N))
    if with_edge_attr and b.get("has_edge_labels", True) and E > 0:
        el = b["elabel"].long()
        emin = int(el.min().item())
        nel = int(num_edge_labels) if num_edge_labels is not None else int(el.max().item()) - emin + 1
        # summed one-hot attributes of merged duplicates = per-label multiplicity of the (src,dst) pair
        oh = torch.zeros((E, nel), dtype=torch.float32, device=dev)
        oh.scatter_(1, (el - emin).view(-1, 1), 1.0)
        group = (keep_scan[1:].long() - 1)  # sorted position -> output edge
        sorted_items = csr_out.eid.long()
        valid = b["src"].long()[sorted_items] != b["dst"].long()[sorted_items]
        edge_attr = torch.zeros((E2, nel), dtype=torch.float32, device=dev)
        edge_attr.index_add_(0, group[valid], oh[sorted_items[valid]])
        out["edge_attr"] = edge_attr
        out["is_dummy_edge"] = edge_attr[:, 0].bool()  # set_dummy_flags: column num_edge_attributes (=0)
    out.lazy("is_dummy_node", lambda: x[:, n_attr].bool())
    if b.get("max_graph_nodes") is not None:
        out["max_graph_nodes"] = int(b["max_graph_nodes"])
    if "y" in b:
        out["y"] = b["y"]
    return out


def process_model_config(config):
    """max_* bookkeeping of the augmentations (subgraph_isomorphism/train.py:38-81)."""
    mc = deepcopy(config)
    if config.get("add_rev", False):
        for k in ("max_nge", "max_ngel", "max_npe", "max_npel"):
            mc[k] *= 2
    if config.get("add_dummy", False):
        mc["max_nge"] += config["max_ngv"] * 2
        mc["max_npe"] += config["max_npv"] * 2
        mc["max_ngel"] += 2
        mc["max_npel"] += 2
        for k in ("max_ngv", "max_npv", "max_ngvl", "max_npvl"):
            mc[k] += 1
    if config.get("convert_conj", False):
        max_ngv, max_npv = mc["max_ngv"], mc["max_npv"]
        avg_gd = math.ceil(mc["max_nge"] / mc["max_ngv"])
        avg_pd = math.ceil(mc["max_npe"] / mc["max_npv"])
        mc["max_ngv"] = mc["max_nge"]
        mc["max_nge"] = (avg_gd * avg_gd) * max_ngv // 2 - max_ngv
        mc["max_npv"] = mc["max_npe"]
        mc["max_npe"] = (avg_pd * avg_pd) * max_npv // 2 - max_npv
        mc["max_ngvl"] = mc["max_ngel"]
        mc["max_npvl"] = mc["max_npel"]
    return mc
