"""ctypes front-end of oracle/c/transforms.c (TEST INFRASTRUCTURE ONLY).

Every function takes/returns host numpy int32 arrays in the batched layout described in
the C file header and cites the reference lines its C body restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "liborc.so")
_lib = None

_I = ctypes.POINTER(ctypes.c_int)
_F = ctypes.POINTER(ctypes.c_float)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_pyg_coalesce.restype = ctypes.c_int64
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_I)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def tu_add_dummy(b):
    """tu_data_processing.py:186-214 (with_dummy=True).  b: raw TU batch dict."""
    B = int(b["num_graphs"])
    N, E = len(b["vlabel"]), len(b["src"])
    o = dict(
        num_graphs=B,
        node_ptr=np.empty(B + 1, np.int32), edge_ptr=np.empty(B + 1, np.int32),
        src=np.empty(E + 2 * N, np.int32), dst=np.empty(E + 2 * N, np.int32),
        vlabel=np.empty(N + B, np.int32), v_is_dummy=np.empty(N + B, np.int32),
        elabel=np.empty(E + 2 * N, np.int32), e_is_dummy=np.empty(E + 2 * N, np.int32),
    )
    lib().orc_tu_add_dummy(
        B, _p(_i32(b["node_ptr"])), _p(_i32(b["edge_ptr"])), _p(_i32(b["src"])), _p(_i32(b["dst"])),
        _p(_i32(b["vlabel"])), _p(_i32(b["elabel"])),
        _p(o["node_ptr"]), _p(o["edge_ptr"]), _p(o["src"]), _p(o["dst"]),
        _p(o["vlabel"]), _p(o["v_is_dummy"]), _p(o["elabel"]), _p(o["e_is_dummy"]))
    if "vattr" in b:  # ATTR passthrough, dummy gets 0 (line 191)
        va = np.zeros(N + B, np.float32)
        real = np.ones(N + B, bool)
        real[o["node_ptr"][1:] - 1] = False
        va[real] = b["vattr"]
        o["vattr"] = va
    # ID = position within the graph (lines 213-214)
    o["vid"] = (np.arange(N + B, dtype=np.int32) - np.repeat(o["node_ptr"][:-1], np.diff(o["node_ptr"]))).astype(np.int32)
    o["eid"] = (np.arange(E + 2 * N, dtype=np.int32) - np.repeat(o["edge_ptr"][:-1], np.diff(o["edge_ptr"]))).astype(np.int32)
    return o


def _two_phase(fn, B, args):
    node_ptr = np.empty(B + 1, np.int32)
    edge_ptr = np.empty(B + 1, np.int32)
    fn(*args, _p(node_ptr), _p(edge_ptr), None, None, None, None)
    V, E = int(node_ptr[-1]), int(edge_ptr[-1])
    src, dst = np.empty(E, np.int32), np.empty(E, np.int32)
    v_origin, e_shared = np.empty(V, np.int32), np.empty(E, np.int32)
    fn(*args, _p(node_ptr), _p(edge_ptr), _p(src), _p(dst), _p(v_origin), _p(e_shared))
    return node_ptr, edge_ptr, src, dst, v_origin, e_shared


def tu_conjugate(b):
    """tu_data_processing.py:223-338.  b: a TU batch (raw -> LINE_, dummy-augmented -> CONJ_).
    Returns the conjugate batch; vertex attributes come from the original edges
    (lines 238-242), edge attributes from the shared original vertex (322-326)."""
    B = int(b["num_graphs"])
    keep = [_i32(b[k]) for k in ("node_ptr", "edge_ptr", "src", "dst", "vlabel")]
    isd = _i32(b.get("e_is_dummy"))
    args = [B] + [_p(a) for a in keep] + [_p(isd)]
    node_ptr, edge_ptr, src, dst, v_origin, e_shared = _two_phase(lib().orc_tu_conjugate, B, args)
    o = dict(num_graphs=B, node_ptr=node_ptr, edge_ptr=edge_ptr, src=src, dst=dst,
             v_origin=v_origin, e_shared=e_shared)
    o["vlabel"] = _i32(b["elabel"])[v_origin]
    o["elabel"] = _i32(b["vlabel"])[e_shared]
    eid = b["eid"] if "eid" in b else (np.arange(len(b["src"]), dtype=np.int32)
                                       - np.repeat(b["edge_ptr"][:-1], np.diff(b["edge_ptr"])))
    vid = b["vid"] if "vid" in b else (np.arange(len(b["vlabel"]), dtype=np.int32)
                                       - np.repeat(b["node_ptr"][:-1], np.diff(b["node_ptr"])))
    o["vid"] = _i32(eid)[v_origin]
    o["eid"] = _i32(vid)[e_shared]
    if isd is not None:
        o["v_is_dummy"] = isd[v_origin]
        o["e_is_dummy"] = _i32(b["v_is_dummy"])[e_shared]
    if "vattr" in b:
        o["eattr"] = b["vattr"][e_shared]
    return o


def sub_add_dummy(b, max_nv, max_nvl, max_ne, max_nel):
    """train.py:404-474 (GraphAdj branch).  max_* are the PRE-augmentation dataset maxima
    passed at train.py:1322-1334."""
    B = int(b["num_graphs"])
    N, E = len(b["vlabel"]), len(b["src"])
    names_n = ("vid", "vlabel", "v_is_dummy")
    names_e = ("eid", "elabel", "e_is_dummy", "e_is_reversed")
    o = dict(num_graphs=B, node_ptr=np.empty(B + 1, np.int32), edge_ptr=np.empty(B + 1, np.int32),
             src=np.empty(E + 2 * N, np.int32), dst=np.empty(E + 2 * N, np.int32))
    for k in names_n:
        o[k] = np.empty(N + B, np.int32)
    for k in names_e:
        o[k] = np.empty(E + 2 * N, np.int32)
    lib().orc_sub_add_dummy(
        B, _p(_i32(b["node_ptr"])), _p(_i32(b["edge_ptr"])), _p(_i32(b["src"])), _p(_i32(b["dst"])),
        _p(_i32(b["vid"])), _p(_i32(b["vlabel"])), _p(_i32(b["eid"])), _p(_i32(b["elabel"])),
        _p(_i32(b.get("e_is_reversed"))),
        int(max_nv), int(max_nvl), int(max_ne), int(max_nel),
        _p(o["node_ptr"]), _p(o["edge_ptr"]), _p(o["src"]), _p(o["dst"]),
        _p(o["vid"]), _p(o["vlabel"]), _p(o["v_is_dummy"]),
        _p(o["eid"]), _p(o["elabel"]), _p(o["e_is_dummy"]), _p(o["e_is_reversed"]))
    return o


def sub_add_reversed(b, max_ne, max_nel):
    """train.py:291-345 (GraphAdj branch), numpy restatement: per graph, the m edges then their m reversals with
    id = arange(max_ne, max_ne + m), label + max_nel, is_reversed = 1 (originals zero-filled)."""
    if "e_is_reversed" in b:
        return b
    B = int(b["num_graphs"])
    cols = {k: [] for k in ("src", "dst", "eid", "elabel", "e_is_reversed")}
    ep = np.zeros(B + 1, np.int32)
    for g in range(B):
        e0, e1 = int(b["edge_ptr"][g]), int(b["edge_ptr"][g + 1])
        m = e1 - e0
        u, v = _i32(b["src"])[e0:e1], _i32(b["dst"])[e0:e1]
        cols["src"] += [u, v]
        cols["dst"] += [v, u]
        cols["eid"] += [_i32(b["eid"])[e0:e1], np.arange(max_ne, max_ne + m, dtype=np.int32)]
        cols["elabel"] += [_i32(b["elabel"])[e0:e1], _i32(b["elabel"])[e0:e1] + np.int32(max_nel)]
        cols["e_is_reversed"] += [np.zeros(m, np.int32), np.ones(m, np.int32)]
        ep[g + 1] = ep[g] + 2 * m
    o = dict(b)
    o["edge_ptr"] = ep
    for k, parts in cols.items():
        o[k] = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    return o


def sub_remove_loops(b):
    """train.py:270-288 (GraphAdj branch): ``remove_edges(e[u == v])`` per graph, order preserved."""
    keep = _i32(b["src"]) != _i32(b["dst"])
    o = dict(b)
    csum = np.concatenate([[0], np.cumsum(keep)]).astype(np.int32)
    o["edge_ptr"] = csum[_i32(b["edge_ptr"])]
    for k in ("src", "dst", "eid", "elabel", "e_is_dummy", "e_is_reversed", "eattr"):
        if k in b:
            o[k] = np.asarray(b[k])[keep]
    return o


def _degrees(b):
    N = len(b["vlabel"])
    return np.bincount(_i32(b["dst"]), minlength=N), np.bincount(_i32(b["src"]), minlength=N)


def compute_norm(b, self_loop=True):
    """utils/graph.py:11-38 -> (node_norm (N,1) float32, edge_norm (E,1))."""
    ind = _degrees(b)[0].astype(np.float32)
    if self_loop:
        nn_ = np.reciprocal(ind + 1)
    else:
        with np.errstate(divide="ignore"):
            nn_ = np.where(ind == 0, np.float32(1.0), np.reciprocal(ind))
    nn_ = nn_.astype(np.float32).reshape(-1, 1)
    return nn_, nn_[_i32(b["dst"])]


def compute_largest_eigenvalues(b):
    """utils/graph.py:41-71 per graph -> (node_eigenv (B,), edge_eigenv (B,)) float32; 0 where a graph has no edge
    (the reference's ``.max()`` of an empty tensor raises there)."""
    ind, outd = _degrees(b)
    B = int(b["num_graphs"])
    ne, ee = np.zeros(B, np.float32), np.zeros(B, np.float32)
    u, v = _i32(b["src"]), _i32(b["dst"])
    for g in range(B):
        e0, e1 = int(b["edge_ptr"][g]), int(b["edge_ptr"][g + 1])
        if e1 > e0:
            ne[g] = (outd[u[e0:e1]] + ind[v[e0:e1]]).max()
            ee[g] = (ind[u[e0:e1]] + outd[v[e0:e1]]).max()
    return ne, ee


def sub_conjugate(b):
    """utils/graph.py:74-175.  Node/edge attribute names are swapped as at lines 155-165:
    the conjugate's vertices carry the edge columns, its edges the shared vertex's columns."""
    B = int(b["num_graphs"])
    keep = [_i32(b[k]) for k in ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "eid")]
    args = [B] + [_p(a) for a in keep]
    node_ptr, edge_ptr, src, dst, v_origin, e_shared = _two_phase(lib().orc_sub_conjugate, B, args)
    o = dict(num_graphs=B, node_ptr=node_ptr, edge_ptr=edge_ptr, src=src, dst=dst,
             v_origin=v_origin, e_shared=e_shared)
    for ek, vk in (("eid", "vid"), ("elabel", "vlabel"), ("e_is_dummy", "v_is_dummy"),
                   ("e_is_reversed", "v_is_reversed")):
        if ek in b:
            o[vk] = _i32(b[ek])[v_origin]
    for vk, ek in (("vid", "eid"), ("vlabel", "elabel"), ("v_is_dummy", "e_is_dummy")):
        if vk in b:
            o[ek] = _i32(b[vk])[e_shared]
    return o


def pyg_coalesce(src, dst, elabel0=None, num_rel=0):
    """PyG read_tu_data: remove_self_loops + coalesce (sort by (row, col), merge duplicates,
    summing one-hot attributes -> per-label multiplicities)."""
    src, dst = _i32(src), _i32(dst)
    el = _i32(elabel0)
    E = len(src)
    n = lib().orc_pyg_coalesce(ctypes.c_int64(E), _p(src), _p(dst), _p(el), int(num_rel), None, None, None, None)
    o_src, o_dst, o_first = (np.empty(n, np.int32) for _ in range(3))
    o_mult = np.zeros((n, num_rel), np.int32) if num_rel else None
    lib().orc_pyg_coalesce(ctypes.c_int64(E), _p(src), _p(dst), _p(el), int(num_rel),
                           _p(o_src), _p(o_dst), _p(o_first), _p(o_mult))
    return o_src, o_dst, o_first, o_mult


def csr_by_dst(N, src, dst):
    src, dst = _i32(src), _i32(dst)
    row_ptr = np.empty(N + 1, np.int32)
    col, eid = np.empty(len(src), np.int32), np.empty(len(src), np.int32)
    lib().orc_csr_by_dst(ctypes.c_int64(N), ctypes.c_int64(len(src)), _p(src), _p(dst),
                         _p(row_ptr), _p(col), _p(eid))
    return row_ptr, col, eid


def spmm_sum(N, src, dst, x, self_scale=0.0):
    src, dst = _i32(src), _i32(dst)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().orc_spmm_sum_f32(ctypes.c_int64(N), ctypes.c_int64(len(src)), int(x.shape[1]), _p(src), _p(dst),
                           x.ctypes.data_as(_F), ctypes.c_float(self_scale), out.ctypes.data_as(_F))
    return out


# ---------------------------------------------------------------------------------------------
# match-weight targets (subgraph_isomorphism/dataset.py:54-108, 1491-1520), plain Python/numpy
def subiso_node_weights(mats, graph_b):
    """per sample: histogram of the subisomorphism entries over the graph's nodes (dataset.py:55-61)."""
    w = np.zeros(len(graph_b["vlabel"]), np.int64)
    for b, m in enumerate(mats):
        n0 = int(graph_b["node_ptr"][b])
        for row in np.asarray(m).reshape(-1, np.asarray(m).shape[-1]) if np.asarray(m).size else []:
            for gu in row:
                w[n0 + int(gu)] += 1
    return w


def subiso_edge_weights(mats, pattern_b, graph_b):
    """per sample: compute_edgeseq_subisoweights (dataset.py:64-108) on the pattern's edges in edge-id order and the
    graph's edges in (src, dst)-sorted order, scattered back to edge-id order (dataset.py:1506-1518)."""
    W = np.zeros(len(graph_b["src"]), np.int64)
    for b, m in enumerate(mats):
        m = np.asarray(m)
        if m.size == 0:
            continue
        pn0, pe0, pe1 = int(pattern_b["node_ptr"][b]), int(pattern_b["edge_ptr"][b]), int(pattern_b["edge_ptr"][b + 1])
        gn0, ge0, ge1 = int(graph_b["node_ptr"][b]), int(graph_b["edge_ptr"][b]), int(graph_b["edge_ptr"][b + 1])
        p_u, p_v = _i32(pattern_b["src"])[pe0:pe1] - pn0, _i32(pattern_b["dst"])[pe0:pe1] - pn0
        p_el = _i32(pattern_b["elabel"])[pe0:pe1]
        g_u, g_v = _i32(graph_b["src"])[ge0:ge1].astype(np.int64) - gn0, _i32(graph_b["dst"])[ge0:ge1].astype(np.int64) - gn0
        g_el = _i32(graph_b["elabel"])[ge0:ge1]
        order = np.lexsort((np.arange(ge1 - ge0), g_v, g_u))          # stable (src, dst) order
        su, sv, sl = g_u[order], g_v[order], g_el[order]
        runs = {}                                                      # (u, v) -> labels of the LAST run with that key
        i = 0
        while i < len(p_el):
            j = i + 1
            while j < len(p_el) and p_u[j] == p_u[i] and p_v[j] == p_v[i]:
                j += 1
            runs[(int(p_u[i]), int(p_v[i]))] = p_el[i:j]
            i = j
        w = np.zeros(ge1 - ge0, np.int64)
        for row in m:
            for (u, v), els in runs.items():
                gu, gv = int(row[u]), int(row[v])
                for k in np.flatnonzero((su == gu) & (sv == gv)):
                    for e in els:
                        if e == sl[k]:
                            w[k] += 1
        W[ge0 + order] = w
    return W


def subiso_conjugate(mats, pattern_b, graph_b):
    """per sample: get_conjugate_subisomorphisms (utils/graph.py:294-330) then ``g_eid[...]`` (train.py:577-587); returns a
    list of (S_b, m_b) int64 matrices of graph-local edge ids (empty (0, m_b) where there is nothing to map)."""
    out = []
    for b, m in enumerate(mats):
        m = np.asarray(m)
        pn0, pe0, pe1 = int(pattern_b["node_ptr"][b]), int(pattern_b["edge_ptr"][b]), int(pattern_b["edge_ptr"][b + 1])
        gn0, ge0, ge1 = int(graph_b["node_ptr"][b]), int(graph_b["edge_ptr"][b]), int(graph_b["edge_ptr"][b + 1])
        p_len = pe1 - pe0
        if m.size == 0 or p_len == 0:
            out.append(np.zeros((0, p_len), np.int64))
            continue
        p_u, p_v = _i32(pattern_b["src"])[pe0:pe1] - pn0, _i32(pattern_b["dst"])[pe0:pe1] - pn0
        p_el = _i32(pattern_b["elabel"])[pe0:pe1]
        g_u, g_v = _i32(graph_b["src"])[ge0:ge1].astype(np.int64) - gn0, _i32(graph_b["dst"])[ge0:ge1].astype(np.int64) - gn0
        g_el = _i32(graph_b["elabel"])[ge0:ge1]
        order = np.lexsort((np.arange(ge1 - ge0), g_v, g_u))
        su, sv, sl = g_u[order], g_v[order], g_el[order]
        runs = {}                                 # insertion order = first appearance; value = labels of the last run
        i = 0
        while i < p_len:
            j = i + 1
            while j < p_len and p_u[j] == p_u[i] and p_v[j] == p_v[i]:
                j += 1
            runs[(int(p_u[i]), int(p_v[i]))] = p_el[i:j]
            i = j
        conj = np.zeros((len(m), p_len), np.int64)
        for r, row in enumerate(m):
            for q, ((u, v), els) in enumerate(runs.items()):
                gu, gv = int(row[u]), int(row[v])
                for k in np.flatnonzero((su == gu) & (sv == gv)):
                    for e in els:
                        if e == sl[k]:
                            conj[r, q] = k
        out.append(order[conj])                   # g_eid[conj_subisomorphisms[i]]
    return out
