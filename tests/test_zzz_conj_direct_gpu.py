"""GPU: the closed-form CONJ_ builder (transforms.tu_conj_structure -> dn4gl_tu_conj_direct_lens / _fill) against the
general path it short-cuts (tu_add_dummy -> tu_conjugate -> pyg_canonicalize -> two CSR builds, itself bit-exact with the
oracle / the reference, tests/test_transforms_gpu.py): same vertex offsets, same CSR pair (row pointers and the column
lists of every row), same features, on the synthetic shapes and on multigraphs with self loops, repeated edges, isolated
nodes, single-node graphs and edgeless graphs; then the same GIN train step losses through ClassificationPipeline."""
from argparse import Namespace

import numpy as np
import pytest
import torch

from dummynode4graphlearning_b200 import synth
from helpers import nasty_tu_batch

pytestmark = pytest.mark.gpu


def _batches():
    yield synth.tu_batch("proteins", 96, seed=5), 2
    yield synth.tu_batch("mutag", 33, seed=6), 5
    for seed in range(10):
        rng = np.random.default_rng(900 + seed)
        b = nasty_tu_batch(rng, int(rng.integers(1, 7)))
        if int(np.diff(b["node_ptr"]).min()) > 0:
            yield b, int(np.asarray(b["elabel"]).max()) + 1 if len(b["elabel"]) else 2


def _rows(csr):
    n = csr.n_rows
    rp = csr.row_ptr[: n + 1].cpu()
    return rp, csr.col[: int(rp[-1])].cpu()


def test_direct_conj_structure_equals_general_path(device):
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.graph import check_errors
    from dummynode4graphlearning_b200.graph_classification.data import Batch
    n = 0
    for raw, nvl in _batches():
        raw = {k: v for k, v in raw.items() if k != "vattr"}
        hint = T.tu_conjugate_sizes_ex(raw, True)
        assert hint[3]
        conj = T.tu_conjugate(T.tu_add_dummy(T.to_device(raw, device)))
        conj.pop("eattr", None)
        conj["has_edge_labels"] = True
        ref = Batch.from_canonical(T.pyg_canonicalize(conj, nvl, None, node_label_min=0, with_edge_attr=False))
        got = T.tu_conj_structure(T.to_device(dict(raw, conj_sizes=hint), device), nvl, 0)
        torch.cuda.synchronize()
        check_errors()
        rs, gs = ref.structure, got.structure
        assert gs.num_nodes == rs.num_nodes and torch.equal(gs.node_ptr.cpu(), rs.node_ptr.cpu())
        for name in ("csr_in", "csr_out"):
            (rp_r, col_r), (rp_g, col_g) = _rows(getattr(rs, name)), _rows(getattr(gs, name))
            assert torch.equal(rp_r, rp_g), name
            assert torch.equal(col_r, col_g), name
        assert torch.equal(ref.x.cpu(), got.x.cpu())
        assert torch.equal(ref.edge_index.cpu(), got.edge_index.cpu())
        assert torch.equal(ref.batch.cpu(), got.batch.cpu())
        # and the aggregation over both structures
        x = torch.rand((gs.num_nodes, 32), device=device)
        from dummynode4graphlearning_b200 import ops
        assert torch.equal(ops.spmm_sum(x, rs.csr_in, rs.csr_out, 1.0), ops.spmm_sum(x, gs.csr_in, gs.csr_out, 1.0))
        n += 1
    assert n >= 6


def test_pipeline_with_extended_hint_takes_the_direct_path_and_matches(device):
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.graph_classification.models import GIN
    from dummynode4graphlearning_b200.optim import FlatAdam
    from dummynode4graphlearning_b200.pipelines import ClassificationPipeline
    raws = [{k: v for k, v in synth.tu_batch("proteins", 48, seed=s).items() if k != "vattr"} for s in (1, 2)]
    order = [0, 1, 0, 0, 1, 1, 0]
    args = Namespace(num_features=2, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                     additional={"train_eps": True, "num_layers": 3, "aggregation": "sum"}, epochs=1, device=str(device))

    def run(hint_fn):
        torch.manual_seed(0)
        model = GIN(args).to(device)
        pipe = ClassificationPipeline(model, FlatAdam(model.parameters(), lr=0.003), mode="conj", num_node_labels=2,
                                      node_label_min=0, cuda_graphs=True)
        devs = [T.to_device(dict(r, conj_sizes=hint_fn(r)) if hint_fn else r, device) for r in raws]
        out = [float(pipe.step_resident(devs[i]).item()) for i in order]
        torch.cuda.synchronize()
        return out

    ref = run(None)
    for fn in (lambda r: T.tu_conjugate_sizes(r, True), lambda r: T.tu_conjugate_sizes_ex(r, True)):
        got = run(fn)
        for a, b in zip(ref, got):
            assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (ref, got)
