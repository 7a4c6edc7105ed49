"""On-disk layout of the counting data sets <-> flat batch dicts (counting side).

The reference reads one GML file per pattern / graph through ``igraph.read`` and one CSV per pattern with the ground
truth (subgraph_isomorphism/utils/io.py:43-220).  This module reads the same directory trees without igraph and hands
the samples over in the flat block-diagonal layout the GPU augmentation kernels consume
(``transforms.to_device`` -> ``sub_add_dummy`` / ``sub_add_reversed`` / ``sub_conjugate``):

* ``parse_gml`` / ``read_gml_graph``: the GML subset igraph writes and reads (``graph [ directed 1  node [ id .. label .. ]
  edge [ source .. target .. label .. key .. ] ]``).  Semantics restated from python-igraph 0.9.11 (README.md:24):
  vertices are numbered in file order, ``source`` / ``target`` name vertices by their ``id`` attribute, numeric attributes
  are doubles -- hence the ``int(...)`` casts of utils/io.py:52-55, which this reader applies as well;
* ``read_graphs_from_dir`` / ``read_patterns_from_dir`` / ``read_metadata_from_dir`` / ``load_data``: the directory
  conventions, the CSV columns and the train / dev / test split rules of utils/io.py:63-220 (same names, same results);
* ``collate``: a list of ``load_data`` samples -> ``(pattern batch, graph batch, counts, subisomorphisms)`` as
  ``GraphAdjDataset.batchify`` groups them (dataset.py:1604-1636), with NODEID from the file and EDGEID = position
  (the igraph branch of ``Graph.__init__``, dataset.py:1101-1130, never sees an edge ``id``);
* ``write_gml_graph`` / ``write_metadata_csv``: the writer side in igraph's GML layout, so synthetic batches can be
  dumped in the reference's format.

Host-side text I/O only -- no device work happens here.
"""
import ast
import csv
import json
import os
import re
from collections import OrderedDict

import numpy as np

csv.field_size_limit(500 * 1024 * 1024)   # utils/io.py:16: the subisomorphism column of one row can be hundreds of MB

_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|\[|\]|[^\s\[\]"]+')


class GMLError(ValueError):
    pass


def _tokens(text):
    for line in text.splitlines():
        s = line.lstrip()
        if s.startswith("#"):          # GML comment lines
            continue
        for m in _TOKEN.finditer(line):
            yield m.group(0)


def _value(tok):
    if tok[0] == '"':
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        try:
            return float(tok)
        except ValueError:
            raise GMLError("bad GML value %r" % tok)


def _parse_list(it):
    """-> list of (key, value) pairs until the closing bracket; nested lists recurse."""
    out = []
    for key in it:
        if key == "]":
            return out
        if key == "[" or key[0] == '"':
            raise GMLError("GML key expected, got %r" % key)
        try:
            tok = next(it)
        except StopIteration:
            raise GMLError("GML key %r without a value" % key)
        if tok == "[":
            out.append((key, _parse_list(it)))
        elif tok == "]":
            raise GMLError("GML key %r without a value" % key)
        else:
            out.append((key, _value(tok)))
    raise GMLError("unterminated GML list")


def parse_gml(text):
    """-> dict(directed, n, src, dst, vattr {name: list}, eattr {name: list}); vertices in file order, edge endpoints
    resolved through the nodes' ``id`` (igraph GML reader semantics).  Attributes missing on some items are None there."""
    it = _tokens(text)
    top = []
    for key in it:
        if key in ("[", "]") or key[0] == '"':
            raise GMLError("GML key expected, got %r" % key)
        try:
            tok = next(it)
        except StopIteration:
            raise GMLError("GML key %r without a value" % key)
        top.append((key, _parse_list(it) if tok == "[" else _value(tok)))
    graphs = [v for k, v in top if k.lower() == "graph" and isinstance(v, list)]
    if len(graphs) != 1:
        raise GMLError("expected exactly one graph [...] block, found %d" % len(graphs))
    directed, nodes, edges = False, [], []
    for k, v in graphs[0]:
        kl = k.lower()
        if kl == "directed":
            directed = bool(v)
        elif kl == "node" and isinstance(v, list):
            nodes.append(v)
        elif kl == "edge" and isinstance(v, list):
            edges.append(v)
    index, vattr = {}, OrderedDict()
    for i, items in enumerate(nodes):
        for k, v in items:
            if isinstance(v, list):
                continue
            vattr.setdefault(k, [None] * len(nodes))[i] = v
            if k == "id":
                if v in index:
                    raise GMLError("duplicate node id %r" % (v,))
                index[v] = i
    if len(index) != len(nodes):
        raise GMLError("node without id")
    src, dst, eattr = [], [], OrderedDict()
    for j, items in enumerate(edges):
        s = t = None
        for k, v in items:
            if isinstance(v, list):
                continue
            if k == "source":
                s = v
            elif k == "target":
                t = v
            else:
                eattr.setdefault(k, [None] * len(edges))[j] = v
        if s not in index or t not in index:
            raise GMLError("edge %d names an unknown node (%r -> %r)" % (j, s, t))
        src.append(index[s])
        dst.append(index[t])
    return dict(directed=directed, n=len(nodes), src=src, dst=dst, vattr=vattr, eattr=eattr)


def _ints(values, what, path):
    if values is None or any(v is None for v in values):
        raise GMLError("%s: missing %s attribute" % (path, what))
    return np.asarray([int(v) for v in values], dtype=np.int64)      # int() truncates like utils/io.py:52-55


def read_gml_graph(path):
    """one pattern / graph file -> dict(num_nodes, src, dst, vid, vlabel, elabel, ekey) (int64, graph-local), i.e. the
    igraph object of utils/io.py:51-55 flattened: ``vs["id"]``, ``vs["label"]``, ``es["label"]``, ``es["key"]`` and
    ``get_edgelist()`` in edge order."""
    with open(path) as f:
        g = parse_gml(f.read())
    n, m = g["n"], len(g["src"])
    return dict(
        num_nodes=g["n"],
        src=np.asarray(g["src"], dtype=np.int64).reshape(m), dst=np.asarray(g["dst"], dtype=np.int64).reshape(m),
        vid=_ints(g["vattr"].get("id") if n else [], "node id", path),
        vlabel=_ints(g["vattr"].get("label") if n else [], "node label", path),
        elabel=_ints(g["eattr"].get("label") if m else [], "edge label", path),
        ekey=_ints(g["eattr"].get("key") if m else [], "edge key", path),
    )


def write_gml_graph(path, g, creator="dn4gl"):
    """the layout python-igraph's ``Graph.write_gml`` produces for a directed graph with vertex attributes id, label and
    edge attributes label, key."""
    n = int(g["num_nodes"]) if "num_nodes" in g else len(g["vlabel"])
    vid = g["vid"] if "vid" in g else np.arange(n)
    ekey = g["ekey"] if "ekey" in g else np.zeros(len(g["src"]), dtype=np.int64)
    out = ['Creator "%s"' % creator, "Version 1", "graph", "[", "  directed 1"]
    for i in range(n):
        out += ["  node", "  [", "    id %d" % int(vid[i]), "    label %d" % int(g["vlabel"][i]), "  ]"]
    idx = {i: int(vid[i]) for i in range(n)}
    for j in range(len(g["src"])):
        out += ["  edge", "  [", "    source %d" % idx[int(g["src"][j])], "    target %d" % idx[int(g["dst"][j])],
                "    label %d" % int(g["elabel"][j]), "    key %d" % int(ekey[j]), "  ]"]
    out.append("]")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


# ----------------------------------------------------------------------------------------------------------------------
# directory conventions (utils/io.py:19-142)
def get_subdirs(dirpath, leaf_only=True):
    """directories below (and including) dirpath, children before their parent, siblings in directory-listing order;
    leaf_only keeps the directories without sub-directories (same list as utils/io.py:19-29)."""
    out, stack = [], [[dirpath, os.scandir(dirpath), True]]
    while stack:
        frame = stack[-1]
        entry = next(frame[1], None)
        if entry is None:
            stack.pop()
            if frame[2] or not leaf_only:
                out.append(frame[0])
        elif entry.is_dir():
            frame[2] = False
            stack.append([entry.path, os.scandir(entry.path), True])
    return out


def get_files(dirpath):
    """every file below dirpath; a sub-directory's files take the place of the sub-directory in its parent's listing
    (same list as utils/io.py:32-40)."""
    out, stack = [], [os.scandir(dirpath)]
    while stack:
        entry = next(stack[-1], None)
        if entry is None:
            stack.pop()
        elif entry.is_dir():
            stack.append(os.scandir(entry.path))
        else:
            out.append(entry.path)
    return out


def _read_graphs_from_dir(dirpath):
    """utils/io.py:43-60: every ``*.gml`` directly under dirpath, keyed by file stem; the first unreadable file ends the
    scan of that directory (the reference prints the exception and breaks)."""
    graphs = {}
    for filename in os.listdir(dirpath):
        if os.path.isdir(os.path.join(dirpath, filename)):
            continue
        stem, ext = os.path.splitext(os.path.basename(filename))
        if ext != ".gml":
            continue
        try:
            graphs[stem] = read_gml_graph(os.path.join(dirpath, filename))
        except Exception as e:
            print(e)
            break
    return graphs


def _map_dirs(fn, items, num_workers):
    if num_workers is not None and num_workers != 1 and len(items) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(num_workers if num_workers > 0 else os.cpu_count()) as ex:
            return list(ex.map(fn, items))
    return [fn(x) for x in items]


def read_graphs_from_dir(dirpath, num_workers=4):
    """utils/io.py:63-78: {leaf directory name: {stem: graph}}, flattened when the leaf is the directory itself."""
    subdirs = get_subdirs(dirpath)
    graphs = {}
    for subdir, x in zip(subdirs, _map_dirs(_read_graphs_from_dir, subdirs, num_workers)):
        graphs[os.path.basename(subdir)] = x
    dirpath = os.path.basename(dirpath)
    if dirpath in graphs and (dirpath == "graphs" or "G_" not in dirpath):
        graphs.update(graphs.pop(dirpath))
    return graphs


def read_patterns_from_dir(dirpath, num_workers=4):
    """utils/io.py:81-96: all leaves merged into one {stem: pattern}."""
    subdirs = get_subdirs(dirpath)
    patterns = {}
    for x in _map_dirs(_read_graphs_from_dir, subdirs, num_workers):
        patterns.update(x)
    dirpath = os.path.basename(dirpath)
    if dirpath in patterns and (dirpath == "patterns" or "P_" not in dirpath):
        patterns.update(patterns.pop(dirpath))
    return patterns


def _parse_subisomorphisms(text):
    """the reference evaluates the cell (utils/io.py:111); the generator writes a Python / JSON list of lists."""
    try:
        v = json.loads(text)
    except ValueError:
        v = ast.literal_eval(text)
    return np.asarray(v, dtype=np.int64)


def _read_metadata_from_csv(csv_file):
    """utils/io.py:99-115: columns g_id, counts, subisomorphisms; an unreadable file yields what was parsed so far."""
    meta = {}
    try:
        with open(csv_file, "r", newline="") as f:
            reader = csv.reader(f, delimiter=",")
            header = next(reader)
            gid_idx, cnt_idx, iso_idx = header.index("g_id"), header.index("counts"), header.index("subisomorphisms")
            for row in reader:
                meta[row[gid_idx]] = {"counts": int(row[cnt_idx]),
                                      "subisomorphisms": _parse_subisomorphisms(row[iso_idx])}
    except Exception as e:
        print(csv_file, e)
    return meta


def read_metadata_from_dir(dirpath, num_workers=4):
    """utils/io.py:118-142: {pattern id (csv stem): {graph id: {counts, subisomorphisms}}}."""
    files = [f for f in get_files(dirpath) if f.endswith(".csv")]
    meta = {}
    for filename, x in zip(files, _map_dirs(_read_metadata_from_csv, files, num_workers)):
        p_id = os.path.splitext(os.path.basename(filename))[0]
        if p_id not in meta:
            meta[p_id] = x
        else:
            meta[p_id].update(x)
    dirpath = os.path.basename(dirpath)
    if dirpath in meta and dirpath == "metadata":
        meta.update(meta.pop(dirpath))
    return meta


def write_metadata_csv(path, rows):
    """rows: iterable of (g_id, counts, subisomorphisms) -> the CSV ``_read_metadata_from_csv`` parses."""
    with open(path, "w", newline="") as f:
        w = csv.writer(f, delimiter=",")
        w.writerow(["g_id", "counts", "subisomorphisms"])
        for g_id, counts, subiso in rows:
            w.writerow([g_id, int(counts), json.dumps(np.asarray(subiso).tolist())])


def _indices(metadata_dir, name):
    p = os.path.join(metadata_dir, name)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return set(int(x) for x in f)


def load_data(pattern_dir, graph_dir, metadata_dir, num_workers=4):
    """utils/io.py:145-220 -> (OrderedDict(train, dev, test), shared_graph).  Each sample is
    ``{"id": "<p>-<g>", "pattern", "graph", "subisomorphisms", "counts"}``.  Splits: explicit ``train.txt`` / ``dev.txt``
    / ``test.txt`` index files under metadata_dir when present; otherwise by the numeric suffix of the graph id --
    modulo 10 (>1 train, 0 dev, 1 test) when every pattern has its own graphs, modulo 3 (2 train, 0 dev, 1 test) when
    the patterns share the graphs."""
    patterns = read_patterns_from_dir(pattern_dir, num_workers=num_workers)
    graphs = read_graphs_from_dir(graph_dir, num_workers=num_workers)
    meta = read_metadata_from_dir(metadata_dir, num_workers=num_workers)
    chosen = {k: _indices(metadata_dir, k + ".txt") for k in ("train", "dev", "test")}
    data = OrderedDict((k, []) for k in ("train", "dev", "test"))
    shared_graph = True
    for p, pattern in patterns.items():
        own = p in graphs
        if own:
            shared_graph = False
        mod = 10 if own else 3
        default = {"train": lambda r: r > 1, "dev": lambda r: r == 0, "test": lambda r: r == 1}
        for g, graph in (graphs[p] if own else graphs).items():
            x = {"id": "%s-%s" % (p, g), "pattern": pattern, "graph": graph,
                 "subisomorphisms": meta[p][g]["subisomorphisms"], "counts": meta[p][g]["counts"]}
            g_idx = int(g.rsplit("_", 1)[-1])
            for split in ("train", "dev", "test"):
                if chosen[split] is not None:
                    if g_idx in chosen[split]:
                        data[split].append(x)
                elif default[split](g_idx % mod):
                    data[split].append(x)
    return data, shared_graph


# ----------------------------------------------------------------------------------------------------------------------
def _pack(graphs):
    B = len(graphs)
    node_ptr = np.zeros(B + 1, dtype=np.int64)
    edge_ptr = np.zeros(B + 1, dtype=np.int64)
    for i, g in enumerate(graphs):
        node_ptr[i + 1] = node_ptr[i] + g["num_nodes"]
        edge_ptr[i + 1] = edge_ptr[i] + len(g["src"])
    if node_ptr[-1] >= 2 ** 31 or edge_ptr[-1] >= 2 ** 31:
        raise ValueError("batch exceeds the library's int32 index range")

    def cat(parts):
        return (np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)).astype(np.int32)

    return dict(
        num_graphs=B, node_ptr=node_ptr.astype(np.int32), edge_ptr=edge_ptr.astype(np.int32),
        src=cat([g["src"] + node_ptr[i] for i, g in enumerate(graphs)]),
        dst=cat([g["dst"] + node_ptr[i] for i, g in enumerate(graphs)]),
        vid=cat([g["vid"] for g in graphs]), vlabel=cat([g["vlabel"] for g in graphs]),
        eid=cat([np.arange(len(g["src"]), dtype=np.int64) for g in graphs]),     # dataset.py:1126-1127: EDGEID = position
        elabel=cat([g["elabel"] for g in graphs]),
    )


def collate(samples):
    """list of ``load_data`` samples -> (pattern batch, graph batch, counts (B,) int64, [subisomorphism matrices]) in the
    layout of ``synth.counting_batch`` / ``matching.pack_subisomorphisms`` (block-diagonal, int32 indices)."""
    counts = np.asarray([int(x["counts"]) for x in samples], dtype=np.int64)
    mats = []
    for x in samples:
        m = np.asarray(x["subisomorphisms"], dtype=np.int64)
        mats.append(m.reshape(-1, x["pattern"]["num_nodes"]) if m.ndim != 2 else m)
    return _pack([x["pattern"] for x in samples]), _pack([x["graph"] for x in samples]), counts, mats
