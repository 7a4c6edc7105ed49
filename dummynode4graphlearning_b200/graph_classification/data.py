"""PyG-style mini-batch (``x, edge_index, edge_attr, batch, y, is_dummy_node, is_dummy_edge``) whose graph
structure is compiled once into the int32 CSR pair the kernels consume.

Stands in for ``torch_geometric.data.Batch`` as produced by PyG's ``DataLoader`` collate
(graph_classification/graph_neural_networks/main.py:245) over ``PYGDataset`` items (dataset.py:118-139).
"""
import torch

from ..graph import build_csr


class GraphStructure:
    """CSR by destination / by source + per-graph node offsets of a PyG-style batch."""

    def __init__(self, edge_index, num_nodes, batch=None, node_ptr=None, max_seg=None, src=None, dst=None,
                 sorted_by_src=False, trash_row=False):
        """edge_index (2, E) int64 as PyG carries it, or the int32 endpoint arrays src / dst directly."""
        if src is None:
            src = edge_index[0].to(torch.int32).contiguous()
            dst = edge_index[1].to(torch.int32).contiguous()
        self.src, self.dst, self.num_nodes = src, dst, int(num_nodes)
        self.trash_row = bool(trash_row)   # src / dst are padded with (N, N) entries up to their capacity (deferred count)
        self.csr_in = build_csr(dst, src, self.num_nodes, trash_row=trash_row)
        self.csr_out = build_csr(src, dst, self.num_nodes, sorted_keys=sorted_by_src, trash_row=trash_row)   # coalesced lists are (src, dst)-sorted
        if node_ptr is None and batch is not None:
            nb = int(batch[-1].item()) + 1 if batch.numel() else 0   # batch is sorted (PyG collate)
            counts = torch.bincount(batch, minlength=nb)
            node_ptr = torch.zeros(nb + 1, dtype=torch.int32, device=batch.device)
            node_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        self.node_ptr = None if node_ptr is None else node_ptr.to(torch.int32).contiguous()
        self.csr_in.seg_ptr = self.csr_out.seg_ptr = self.node_ptr
        self.csr_in.max_seg = self.csr_out.max_seg = max_seg   # rows of the largest graph, if the host knows it
        self._rel = {}
        self._row2seg = None

    @classmethod
    def from_csr(cls, csr_in, csr_out, num_nodes, node_ptr, max_seg=None):
        """from an already compiled CSR pair (transforms.tu_conj_structure): no edge list is kept (src / dst are None)."""
        s = cls.__new__(cls)
        s.src = s.dst = None
        s.num_nodes, s.trash_row = int(num_nodes), False
        s.csr_in, s.csr_out = csr_in, csr_out
        s.node_ptr = node_ptr.to(torch.int32).contiguous()
        s.csr_in.seg_ptr = s.csr_out.seg_ptr = s.node_ptr
        s.csr_in.max_seg = s.csr_out.max_seg = max_seg
        s._rel, s._row2seg = {}, None
        return s

    @property
    def row2seg(self):
        """int32 row -> graph index (``batch`` as the kernels read it); built once per batch."""
        if self._row2seg is None and self.node_ptr is not None:
            from .. import ops
            self._row2seg = ops.segment_ids(self.node_ptr, self.num_nodes)
        return self._row2seg

    def relation_csr(self, edge_type, num_rels):
        """CSR pair addressing the (N*R, D) per-relation table (see subgraph_isomorphism/models/rgin.py)."""
        from ..graph import CSR, cached_for_tensor

        def make():
            base = self.csr_in
            et = edge_type.long()
            col = (base.col.long() * num_rels + et[base.eid.long()]).to(torch.int32)
            fwd = CSR(base.row_ptr, col, base.eid, base.n_rows, base.nnz)
            fwd.heavy_rows, fwd.heavy_count, fwd.heavy_thr = base.heavy_rows, base.heavy_count, base.heavy_thr
            bwd = build_csr((self.src.long() * num_rels + et).to(torch.int32), self.dst, self.num_nodes * num_rels)
            return fwd, bwd

        return cached_for_tensor(self._rel, ("rel_csr", int(num_rels)), edge_type, int(num_rels), make)


class Batch:
    """``edge_index`` / ``batch`` / ``is_dummy_node`` may be given as zero-argument callables: they are then built on
    first access (the GIN train step never reads them -- it uses the compiled structure)."""

    def __init__(self, x, edge_index, batch, edge_attr=None, y=None, is_dummy_node=None, is_dummy_edge=None,
                 node_ptr=None, max_graph_nodes=None, src=None, dst=None, sorted_by_src=False, trash_row=False):
        self.x, self.edge_attr, self.y = x, edge_attr, y
        self._sorted_by_src, self._trash_row = sorted_by_src, trash_row
        self._lazy = dict(edge_index=edge_index, batch=batch, is_dummy_node=is_dummy_node)
        self.is_dummy_edge = is_dummy_edge
        self._node_ptr, self._max_seg = node_ptr, max_graph_nodes
        self._src, self._dst = src, dst
        self._structure = None

    def _get(self, name):
        v = self._lazy[name]
        if callable(v):
            v = self._lazy[name] = v()
        return v

    edge_index = property(lambda self: self._get("edge_index"))
    batch = property(lambda self: self._get("batch"))
    is_dummy_node = property(lambda self: self._get("is_dummy_node"))

    @property
    def num_graphs(self):
        return int(self.structure.node_ptr.numel()) - 1

    @property
    def structure(self):
        if self._structure is None:
            if self._src is not None:
                self._structure = GraphStructure(None, self.x.size(0), None if self._node_ptr is not None else self.batch,
                                                 self._node_ptr, self._max_seg, src=self._src, dst=self._dst,
                                                 sorted_by_src=self._sorted_by_src, trash_row=self._trash_row)
            else:
                self._structure = GraphStructure(self.edge_index, self.x.size(0), self.batch, self._node_ptr, self._max_seg)
        return self._structure

    @staticmethod
    def from_canonical(d):
        """from ``transforms.pyg_canonicalize`` output."""
        lazy = lambda k: (lambda: d[k]) if k in d else None
        return Batch(d["x"], lazy("edge_index"), lazy("batch"), d.get("edge_attr"), d.get("y"),
                     lazy("is_dummy_node"), d.get("is_dummy_edge"), node_ptr=d["node_ptr"],
                     max_graph_nodes=d.get("max_graph_nodes"), src=d.get("src"), dst=d.get("dst"),
                     sorted_by_src=bool(d.get("sorted_by_src", False)), trash_row=bool(d.get("trash_row", False)))


def structure_of(data):
    """GraphStructure of any object with ``x, edge_index, batch`` (cached on the object)."""
    s = getattr(data, "_structure", None)
    if s is None:
        s = GraphStructure(data.edge_index, data.x.size(0), data.batch, getattr(data, "_node_ptr", None),
                           getattr(data, "_max_seg", None))
        try:
            data._structure = s
        except AttributeError:
            pass
    return s
