"""Minimal list-backed stand-in for python-igraph 0.9 (TEST SCAFFOLDING, not product code).

Implements just the container semantics the reference relies on
(graph_classification/data_processing/tu_data_processing.py:185-214, 223-338 and
subgraph_isomorphism/utils/graph.py:177-267): ordered vertex/edge attribute lists,
``incident(v, mode)``, ``get_edgelist()`` in edge-id order, and ``delete_vertices`` that
compacts the survivors preserving relative order (igraph semantics).

An in-edge index is maintained so ``incident`` is O(deg) instead of O(E); results are the
same ascending edge ids igraph returns.
"""


class _Attr:
    __slots__ = ("g", "kind", "i")

    def __init__(self, g, kind, i):
        self.g, self.kind, self.i = g, kind, i

    @property
    def source(self):
        return self.g._edges[self.i][0]

    @property
    def target(self):
        return self.g._edges[self.i][1]

    @property
    def index(self):
        return self.i

    def __getitem__(self, k):
        store = self.g._vattr if self.kind == "v" else self.g._eattr
        return store[k][self.i]


class _Seq:
    __slots__ = ("g", "kind")

    def __init__(self, g, kind):
        self.g, self.kind = g, kind

    def _store(self):
        return self.g._vattr if self.kind == "v" else self.g._eattr

    def __len__(self):
        return self.g._n if self.kind == "v" else len(self.g._edges)

    def __setitem__(self, k, v):
        v = list(v)
        if len(v) != len(self):
            raise ValueError("attribute list length %d != %d" % (len(v), len(self)))
        self._store()[k] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return self._store()[k]
        return _Attr(self.g, self.kind, k)

    def __delitem__(self, k):
        del self._store()[k]


class Graph:
    def __init__(self, directed=True):
        self._directed = directed
        self._n = 0
        self._edges = []
        self._vattr = {}
        self._eattr = {}
        self._in = []
        self._out = []

    vs = property(lambda s: _Seq(s, "v"))
    es = property(lambda s: _Seq(s, "e"))

    def add_vertices(self, n):
        self._n += n
        self._in.extend([] for _ in range(n))
        self._out.extend([] for _ in range(n))
        for k in self._vattr:
            self._vattr[k].extend([None] * n)

    def add_edges(self, es):
        cnt = 0
        for a, b in es:
            a, b = int(a), int(b)
            eid = len(self._edges)
            self._edges.append((a, b))
            self._out[a].append(eid)
            self._in[b].append(eid)
            cnt += 1
        for k in self._eattr:
            self._eattr[k].extend([None] * cnt)

    def vcount(self):
        return self._n

    def ecount(self):
        return len(self._edges)

    def vertex_attributes(self):
        return list(self._vattr)

    def edge_attributes(self):
        return list(self._eattr)

    def incident(self, v, mode="out"):
        mode = str(mode).lower()
        if mode == "in":
            return list(self._in[v])
        if mode == "out":
            return list(self._out[v])
        return sorted(self._in[v] + self._out[v])

    def indegree(self):
        return [len(x) for x in self._in]

    def outdegree(self):
        return [len(x) for x in self._out]

    def get_edgelist(self):
        return list(self._edges)

    def delete_vertices(self, vs):
        dead = set(int(v) for v in vs)
        keep = [v for v in range(self._n) if v not in dead]
        remap = {v: i for i, v in enumerate(keep)}
        keep_e = [i for i, (a, b) in enumerate(self._edges) if a in remap and b in remap]
        new_edges = [(remap[self._edges[i][0]], remap[self._edges[i][1]]) for i in keep_e]
        for k in self._eattr:
            self._eattr[k] = [self._eattr[k][i] for i in keep_e]
        for k in self._vattr:
            self._vattr[k] = [self._vattr[k][v] for v in keep]
        self._n = len(keep)
        self._edges = []
        self._in = [[] for _ in range(self._n)]
        self._out = [[] for _ in range(self._n)]
        for eid, (a, b) in enumerate(new_edges):
            self._edges.append((a, b))
            self._out[a].append(eid)
            self._in[b].append(eid)


# utils/graph.py:177 dispatches on str(graph.__class__) == "<class 'igraph.Graph'>"
Graph.__module__ = "igraph"


def read(path, *a, **k):
    """``igraph.read`` for the GML files of the counting data sets (utils/io.py:51).  Line-oriented restatement of the
    python-igraph 0.9 GML reader for the layout igraph itself writes (one ``key value`` per line, ``[`` / ``]`` on their own
    line or after the key): vertices numbered in file order, ``source`` / ``target`` resolved through the nodes' ``id``,
    numeric attributes returned as floats (igraph's numeric attribute type is double)."""
    directed, nodes, edges, cur, kind = False, [], [], None, None
    pending = None
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("#"):
                continue
            parts = line.split(None, 1)
            key = parts[0]
            val = parts[1].strip() if len(parts) > 1 else None
            if key == "[":
                if pending in ("node", "edge"):
                    kind, cur = pending, {}
                pending = None
                continue
            if key == "]":
                if kind == "node":
                    nodes.append(cur)
                elif kind == "edge":
                    edges.append(cur)
                kind, cur = None, None
                continue
            if key in ("graph", "node", "edge") and val in (None, "["):
                if val == "[" and key in ("node", "edge"):
                    kind, cur = key, {}
                else:
                    pending = key
                continue
            if kind is None:
                if key == "directed":
                    directed = bool(int(val))
                continue                       # Creator / Version / graph attributes
            cur[key] = val[1:-1] if val.startswith('"') else float(val)
    g = Graph(directed=directed)
    g.add_vertices(len(nodes))
    index = {nd["id"]: i for i, nd in enumerate(nodes)}
    for key in dict.fromkeys(k2 for nd in nodes for k2 in nd):
        g.vs[key] = [nd.get(key) for nd in nodes]
    g.add_edges([(index[e["source"]], index[e["target"]]) for e in edges])
    for key in dict.fromkeys(k2 for e in edges for k2 in e if k2 not in ("source", "target")):
        g.es[key] = [e.get(key) for e in edges]
    return g
