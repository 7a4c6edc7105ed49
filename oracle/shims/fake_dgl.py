"""Tiny stand-in for DGL >= 0.6 (TEST SCAFFOLDING, not product code).

Restates the documented DGL semantics the reference's hot path depends on (SURVEY.md
App. C): frames are dicts of tensors; ``update_all(udf, fn.sum(msg,out), apply)`` calls
the message UDF once over all edges in eid order, sum-reduces the named message onto
destination nodes (zeros for isolated nodes) and merges the apply UDF's dict into ndata;
``apply_edges`` merges into edata; side-effect writes to ``edges.data`` inside a message
UDF persist (subgraph_isomorphism/models/dmpnn.py:126 relies on it); ``add_nodes`` /
``add_edges`` zero-fill frame columns that the caller did not provide and create new
columns zero-filled for pre-existing rows; ``remove_nodes`` compacts survivors keeping
relative order; ``dgl.batch`` is the block-diagonal union with frames concatenated.

Only ``fn.sum`` (the gspmm copy_e/sum reduce) is restated arithmetic: ``index_add_`` in
eid order on a zeros tensor.
"""
import torch as th


class _SumReduce:
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


class _TargetCode:      # dgl 0.4 ``fn.TargetCode`` (rgcn.py:157)
    SRC, DST, EDGE = 0, 1, 2


class function:  # stands for the module ``dgl.function``
    TargetCode = _TargetCode

    @staticmethod
    def sum(msg, out):
        return _SumReduce(msg, out)

    @staticmethod
    def CopyMessageFunction(target, in_field, out_field):
        """dgl 0.4 builtin: copy a source / destination node field (or an edge field) onto the edges."""
        def f(edges):
            frame = {_TargetCode.SRC: edges.src, _TargetCode.DST: edges.dst, _TargetCode.EDGE: edges.data}[target]
            return {out_field: frame[in_field]}
        return f

    @staticmethod
    def copy_u(u, out):
        return lambda edges: {out: edges.src[u]}


class _EdgeBatch:
    def __init__(self, g):
        self._g = g
        self.src = {k: v[g._u] for k, v in g.ndata.items()}
        self.dst = {k: v[g._v] for k, v in g.ndata.items()}
        self.data = g.edata  # same dict: UDF side effects persist

    def __len__(self):
        return self._g.number_of_edges()


class _NodeBatch:
    def __init__(self, g):
        self.data = g.ndata


def _zero_rows(like, n):
    return th.zeros((n,) + tuple(like.shape[1:]), dtype=like.dtype, device=like.device)


class DGLGraph:
    def __init__(self, u=None, v=None, num_nodes=0, batch_num_nodes=None, batch_num_edges=None, **kw):
        self._u = th.zeros((0,), dtype=th.long) if u is None else th.as_tensor(u, dtype=th.long)
        self._v = th.zeros((0,), dtype=th.long) if v is None else th.as_tensor(v, dtype=th.long)
        self._n = int(num_nodes)
        self.ndata = {}
        self.edata = {}
        self._bnn = batch_num_nodes
        self._bne = batch_num_edges

    # ---- structure queries -------------------------------------------------------------
    @property
    def batch_size(self):
        return 1 if self._bnn is None else len(self._bnn)

    def batch_num_nodes(self):
        return th.tensor([self._n]) if self._bnn is None else self._bnn

    def batch_num_edges(self):
        return th.tensor([len(self._u)]) if self._bne is None else self._bne

    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._u.numel())

    def in_degrees(self):
        return th.bincount(self._v, minlength=self._n)

    def out_degrees(self):
        return th.bincount(self._u, minlength=self._n)

    def all_edges(self, form="uv", order="eid"):
        if order == "eid":
            e = th.arange(len(self._u))
        elif order == "srcdst":
            key = self._u * max(self._n, 1) + self._v
            e = th.sort(key, stable=True)[1]
        else:
            raise ValueError(order)
        if form == "uv":
            return self._u[e], self._v[e]
        if form == "all":
            return self._u[e], self._v[e], e
        if form == "eid":
            return e
        raise ValueError(form)

    edges = all_edges

    def incidence_matrix(self, typestr):
        assert typestr == "in"
        E = self.number_of_edges()
        idx = th.stack([self._v, th.arange(E)])
        return th.sparse_coo_tensor(idx, th.ones(E), (self._n, E)).coalesce()

    # ---- mutation ------------------------------------------------------------------------
    def _grow(self, frame, old_rows, new_rows, data):
        data = dict(data or {})
        for k, val in data.items():
            if k not in frame:
                frame[k] = _zero_rows(val, old_rows)
        for k in list(frame.keys()):
            add = data[k] if k in data else _zero_rows(frame[k], new_rows)
            frame[k] = th.cat([frame[k], add.to(frame[k].dtype)], dim=0)

    def add_nodes(self, num, data=None):
        self._grow(self.ndata, self._n, num, data)
        self._n += num

    def add_edges(self, u, v, data=None):
        u = th.as_tensor(u, dtype=th.long).view(-1)
        v = th.as_tensor(v, dtype=th.long).view(-1)
        self._grow(self.edata, self.number_of_edges(), len(u), data)
        self._u = th.cat([self._u, u])
        self._v = th.cat([self._v, v])

    def remove_edges(self, eids):
        """DGL semantics: the listed edges disappear, the others keep their relative order and attribute rows."""
        dead = th.zeros(self.number_of_edges(), dtype=th.bool)
        dead[th.as_tensor(eids, dtype=th.long).view(-1)] = True
        keep = (~dead).nonzero().view(-1)
        self._u, self._v = self._u[keep], self._v[keep]
        for k in self.edata:
            self.edata[k] = self.edata[k][keep]

    def remove_nodes(self, nids):
        dead = th.zeros(self._n, dtype=th.bool)
        dead[th.as_tensor(nids, dtype=th.long)] = True
        keep_n = (~dead).nonzero().view(-1)
        remap = th.full((self._n,), -1, dtype=th.long)
        remap[keep_n] = th.arange(len(keep_n))
        keep_e = (~(dead[self._u] | dead[self._v])).nonzero().view(-1)
        self._u, self._v = remap[self._u[keep_e]], remap[self._v[keep_e]]
        for k in self.ndata:
            self.ndata[k] = self.ndata[k][keep_n]
        for k in self.edata:
            self.edata[k] = self.edata[k][keep_e]
        self._n = len(keep_n)

    # ---- message passing ------------------------------------------------------------------
    def update_all(self, message_func, reduce_func, apply_node_func=None):
        msgs = message_func(_EdgeBatch(self))
        m = msgs[reduce_func.msg]
        out = th.zeros((self._n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        self.ndata[reduce_func.out] = out.index_add_(0, self._v.to(m.device), m)
        if apply_node_func is not None:
            self.ndata.update(apply_node_func(_NodeBatch(self)))

    def apply_edges(self, func):
        self.edata.update(func(_EdgeBatch(self)))

    def to(self, device):
        return self

    def local_var(self):
        return self


# utils/graph.py:77-81 dispatches on the class name string
DGLGraph.__module__ = "dgl.graph"

__version__ = "0.6.1"


def batch(graphs):
    """Block-diagonal union (SURVEY.md App. C ``dgl.batch``)."""
    us, vs, off = [], [], 0
    for g in graphs:
        us.append(g._u + off)
        vs.append(g._v + off)
        off += g._n
    out = DGLGraph(
        th.cat(us), th.cat(vs), off,
        batch_num_nodes=th.tensor([g._n for g in graphs], dtype=th.long),
        batch_num_edges=th.tensor([g.number_of_edges() for g in graphs], dtype=th.long),
    )
    for k in graphs[0].ndata:
        out.ndata[k] = th.cat([g.ndata[k] for g in graphs], dim=0)
    for k in graphs[0].edata:
        out.edata[k] = th.cat([g.edata[k] for g in graphs], dim=0)
    return out
