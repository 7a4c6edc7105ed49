"""On-disk TU layout <-> flat batch dicts.

``load_tu_dir`` parses what ``load_graph_data_from_TUDatadir`` parses
(graph_classification/data_processing/tu_data_processing.py:125-185: ``*_A.txt`` with 1-based global node ids,
``*_graph_indicator.txt``, ``*_node_labels.txt``, ``*_edge_labels.txt``, ``*_node_attributes.txt``,
``*_edge_attributes.txt``; labels shifted so the smallest is 1, missing label files -> all ones) into the flat
block-diagonal batch the GPU transforms consume (``transforms.to_device`` -> ``tu_add_dummy`` -> ``tu_conjugate``).
``save_tu_dir`` writes a (transformed) batch back exactly as ``save_graph_data`` / ``save_graph_labels`` do
(tu_data_processing.py:341-414), so DUMMY_/LINE_/CONJ_ datasets produced on the GPU are byte-identical to the
reference's offline files and can be read by its ``PYGDataset`` / PyG ``read_tu_data``.
Host-side text I/O only -- no device work happens here.
"""
import os

import numpy as np


def _read_ints(path):
    with open(path) as f:
        return np.array([int(line.strip()) for line in f if line.strip() != ""], dtype=np.int64)


def _read_floats(path):
    with open(path) as f:
        return [float(line.strip()) for line in f if line.strip() != ""]


def _shift_min_to_one(labels):
    """tu_data_processing.py:157-171"""
    m = int(labels.min())
    return labels - m + 1 if m != 1 else labels


def load_tu_dir(data_dir, keep_trailing_edgeless=False):
    """-> dict(num_graphs, node_ptr, edge_ptr, src, dst, vlabel, elabel[, vattr][, eattr][, y], has_edge_labels)
    with graph-local structure flattened to global 0-based node ids (int32).

    Like the reference's walk over the edge list (``while j < k``, tu_data_processing.py:179), graphs that come AFTER
    the graph of the last edge are not produced (trailing graphs without edges are dropped; a data
    set without any edge yields no graph).  ``y`` is cut to the graphs produced, all labels of the file stay available
    as ``y_all`` (the reference saves them unchanged, :341-350).  keep_trailing_edgeless=True keeps every graph."""
    if os.path.exists(os.path.join(data_dir, "raw")):
        data_dir = os.path.join(data_dir, "raw")
    A, gi, nl, el, na, ea, y = [], [], [], [], [], [], None
    for fn in sorted(os.listdir(data_dir)):
        p = os.path.join(data_dir, fn)
        if fn.endswith("_A.txt"):
            with open(p) as f:
                A.extend(tuple(map(int, line.strip().replace(" ", "").split(","))) for line in f if line.strip() != "")
        elif fn.endswith("_graph_indicator.txt"):
            gi.extend(_read_ints(p).tolist())
        elif fn.endswith("_node_labels.txt"):
            nl.extend(_read_ints(p).tolist())
        elif fn.endswith("_edge_labels.txt"):
            el.extend(_read_ints(p).tolist())
        elif fn.endswith("_node_attributes.txt"):
            na.extend(_read_floats(p))
        elif fn.endswith("_edge_attributes.txt"):
            ea.extend(_read_floats(p))
        elif fn.endswith("_graph_labels.txt"):
            y = _read_ints(p)
    gi = np.asarray(gi, dtype=np.int64)
    A = np.asarray(A, dtype=np.int64).reshape(-1, 2)
    N, E = len(gi), len(A)
    has_el = len(el) > 0
    vlabel = _shift_min_to_one(np.asarray(nl, dtype=np.int64)) if len(nl) else np.ones(N, np.int64)
    elabel = _shift_min_to_one(np.asarray(el, dtype=np.int64)) if has_el else np.ones(E, np.int64)
    # graphs are numbered 1..B in file order; nodes of a graph are contiguous (the reference walks them that way, :173-185)
    B = int(gi.max()) if N else 0
    counts = np.bincount(gi - 1, minlength=B)
    node_ptr = np.concatenate([[0], np.cumsum(counts)])
    src, dst = A[:, 0] - 1, A[:, 1] - 1
    eg = gi[src] - 1 if E else np.zeros(0, np.int64)
    if E and (np.any(np.diff(eg) < 0) or np.any(gi[dst] - 1 != eg)):
        raise ValueError("load_tu_dir: edges must be grouped by graph and stay inside their graph (TU layout)")
    edge_ptr = np.concatenate([[0], np.cumsum(np.bincount(eg, minlength=B))])
    out = dict(num_graphs=B, node_ptr=node_ptr.astype(np.int32), edge_ptr=edge_ptr.astype(np.int32),
               src=src.astype(np.int32), dst=dst.astype(np.int32), vlabel=vlabel.astype(np.int32),
               elabel=elabel.astype(np.int32), has_edge_labels=has_el)
    if na:
        out["vattr"] = np.asarray(na, dtype=np.float64)
    if ea:
        out["eattr"] = np.asarray(ea, dtype=np.float64)
    if y is not None:
        out["y"] = y
    if not keep_trailing_edgeless:
        B_eff = int(eg[-1]) + 1 if E else 0
        if B_eff < B:
            n_eff = int(node_ptr[B_eff])
            out.update(num_graphs=B_eff, node_ptr=out["node_ptr"][:B_eff + 1], edge_ptr=out["edge_ptr"][:B_eff + 1],
                       vlabel=out["vlabel"][:n_eff])
            if "vattr" in out:
                out["vattr"] = out["vattr"][:n_eff]
            if y is not None:
                out["y_all"], out["y"] = y, y[:B_eff]
    return out


def _np(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)


def tu_file_lines(b):
    """suffix -> list of text lines, exactly what save_graph_data / save_graph_labels write (tu_data_processing.py:341-414)."""
    node_ptr, src, dst = _np(b["node_ptr"]).astype(np.int64), _np(b["src"]).astype(np.int64), _np(b["dst"]).astype(np.int64)
    B = int(b["num_graphs"])
    res = {
        "graph_indicator": [str(g + 1) for g in range(B) for _ in range(int(node_ptr[g + 1] - node_ptr[g]))],
        "A": ["%d,%d" % (s + 1, d + 1) for s, d in zip(src, dst)],            # global ids, 1-based (graph_nsum starts at 1)
        "node_labels": [str(int(x)) for x in _np(b["vlabel"])],
        "edge_labels": [str(int(x)) for x in _np(b["elabel"])],
    }
    if "vattr" in b:
        res["node_attributes"] = [str(float(x)) for x in _np(b["vattr"])]
    if "eattr" in b:
        res["edge_attributes"] = [str(float(x)) for x in _np(b["eattr"])]
    res["node_ids"] = [str(int(x)) for x in _np(b["vid"])]
    res["edge_ids"] = [str(int(x)) for x in _np(b["eid"])]
    if "y" in b:
        res["graph_labels"] = [str(int(x)) for x in _np(b["y"])]
    return res


def save_tu_dir(b, data_dir, prefix=""):
    """writes ``<prefix>{graph_indicator,A,node_labels,edge_labels,[node_attributes],[edge_attributes],node_ids,
    edge_ids,[graph_labels]}.txt`` into data_dir; default prefix as in the reference (directory name + '_', or the
    parent's name for a ``raw`` directory)."""
    if prefix == "":
        prefix = os.path.basename(data_dir) + "_"
        if prefix == "raw_":
            prefix = os.path.basename(os.path.dirname(data_dir)) + "_"
    os.makedirs(data_dir, exist_ok=True)
    for suffix, lines in tu_file_lines(b).items():
        with open(os.path.join(data_dir, prefix + suffix + ".txt"), "w") as f:
            for line in lines:
                f.write(line)
                f.write("\n")
    return data_dir


def convert_tu_dataset(raw_path, dataset, device="cuda:0"):
    """what running ``tu_data_processing.py`` does after its download step (:431-456): read ``<...>/<dataset>/raw``,
    build the DUMMY_ (dummy-augmented), LINE_ (edge-to-vertex) and CONJ_ (dummy + edge-to-vertex) variants and write
    them next to it as ``.../DUMMY_<dataset>/raw`` etc., graph labels included -- with the graph construction done by
    the GPU kernels on the whole data set at once instead of igraph calls per graph.  Scalar node / edge attributes
    never leave the host: they are carried in float64 through the index maps the kernels return (dummy items get 0,
    :191,:197; conjugate vertices take their original edge's attributes and conjugate edges their shared vertex's,
    :238-242,:322-326), so the text written equals the reference's digit for digit.  Returns {variant: directory}."""
    from .. import transforms as T
    raw = load_tu_dir(raw_path)
    dev = T.to_device({k: v for k, v in raw.items() if k not in ("y", "vattr", "eattr")}, device)
    dummy = T.tu_add_dummy(dev)
    line, conj = T.tu_conjugate(dev), T.tu_conjugate(dummy)

    def with_dummy_zeros(values, flags):
        out = np.zeros(len(flags), dtype=np.float64)
        out[flags == 0] = values
        return out

    attrs = {"DUMMY_": {}, "LINE_": {}, "CONJ_": {}}
    va, ea = raw.get("vattr"), raw.get("eattr")
    va_d = None if va is None else with_dummy_zeros(va, _np(dummy["v_is_dummy"]))
    ea_d = None if ea is None else with_dummy_zeros(ea, _np(dummy["e_is_dummy"]))
    for key, val in (("vattr", va_d), ("eattr", ea_d)):
        if val is not None:
            attrs["DUMMY_"][key] = val
    for prefix, b, v_src, e_src in (("LINE_", line, va, ea), ("CONJ_", conj, va_d, ea_d)):
        if e_src is not None:
            attrs[prefix]["vattr"] = e_src[_np(b["v_origin"]).astype(np.int64)]
        if v_src is not None:
            attrs[prefix]["eattr"] = v_src[_np(b["e_shared"]).astype(np.int64)]
    out = {}
    for prefix, b in (("DUMMY_", dummy), ("LINE_", line), ("CONJ_", conj)):
        target = raw_path.replace(dataset, prefix + dataset)       # tu_data_processing.py:438-440
        if target == raw_path:
            raise ValueError("convert_tu_dataset: %r does not occur in %r" % (dataset, raw_path))
        b = {k: b[k] for k in ("num_graphs", "node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "vid", "eid")}
        b.update(attrs[prefix])
        if "y" in raw:
            b["y"] = raw.get("y_all", raw["y"])                    # save_graph_labels writes the label file unchanged
        out[prefix] = save_tu_dir(b, target)
    return out
