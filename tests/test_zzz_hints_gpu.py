"""GPU: the host-side size hint of the edge-to-vertex transform (transforms.tu_conjugate_sizes -> batch["conj_sizes"]):
with it tu_conjugate allocates its outputs without reading the sizes back; outputs must equal the un-hinted call and the
oracle bit for bit, and a hint that does not match what the device counted must raise the asynchronous error flag.
(Written after the round's GPU budget was spent: no B200 log yet, hence the test_zzz_ prefix.)"""
import numpy as np
import pytest
import torch

from dummynode4graphlearning_b200 import synth
from helpers import nasty_tu_batch
from oracle import transforms as OT

pytestmark = pytest.mark.gpu

KEYS = ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "vid", "eid", "v_is_dummy", "e_is_dummy")


def _batches():
    yield synth.tu_batch("proteins", 64, seed=5)
    yield synth.tu_batch("mutag", 33, seed=6)
    for seed in range(6):
        rng = np.random.default_rng(500 + seed)
        yield nasty_tu_batch(rng, int(rng.integers(1, 6)))


@pytest.mark.parametrize("with_dummy", [True, False])
def test_hinted_conjugate_equals_unhinted_and_oracle(device, with_dummy):
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.graph import check_errors
    for raw in _batches():
        ref = OT.tu_conjugate(OT.tu_add_dummy(raw)) if with_dummy else OT.tu_conjugate(raw)
        hinted = dict(raw, conj_sizes=T.tu_conjugate_sizes(raw, with_dummy))
        outs = []
        for b in (raw, hinted):
            dev = T.to_device(b, device)
            outs.append(T.tu_conjugate(T.tu_add_dummy(dev) if with_dummy else dev))
        torch.cuda.synchronize()
        check_errors()
        assert outs[0]["max_graph_nodes"] == outs[1]["max_graph_nodes"]
        for k in KEYS:
            if k in ref:
                want = np.asarray(ref[k]).astype(np.int64)
                for o in outs:
                    assert np.array_equal(o[k].cpu().numpy().astype(np.int64), want), k


def test_wrong_hint_raises_the_async_flag(device):
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.graph import check_errors
    raw = synth.tu_batch("mutag", 8, seed=1)
    v, e, mx = T.tu_conjugate_sizes(raw, True)
    T.tu_conjugate(T.tu_add_dummy(T.to_device(dict(raw, conj_sizes=(v + 1, e + 3, mx)), device)))   # too large: no overrun
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="error -5"):
        check_errors()
    check_errors()     # the flag was cleared


def test_too_small_hint_writes_nothing_out_of_bounds(device):
    """a hint SMALLER than what the device counts: the fill kernel must not write beyond the caller's buffers (capacity
    arguments of dn4gl_tu_conjugate_fill), and the flag is raised."""
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.graph import check_errors
    raw = synth.tu_batch("mutag", 8, seed=1)
    v, e, mx = T.tu_conjugate_sizes(raw, True)
    guard = torch.full((4096,), 12345, dtype=torch.int32, device=device)      # likely neighbours of the small outputs
    out = T.tu_conjugate(T.tu_add_dummy(T.to_device(dict(raw, conj_sizes=(v - 2, e - 50, mx)), device)))
    torch.cuda.synchronize()
    assert out["src"].numel() == e - 50 and bool((guard == 12345).all())
    with pytest.raises(RuntimeError, match="error -5"):
        check_errors()


def test_captured_transform_equals_eager_and_pipeline_trajectories(device):
    """ClassificationPipeline with the loader's size hint: the transform is replayed as a CUDA graph from the third step on
    (outputs are static tensors guarded by the train step's 'copied' event).  Same structure tensors as the eager transform
    and the same loss trajectory as a pipeline without hints, through both entry points (resident and pinned host batch)."""
    from argparse import Namespace
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.graph_classification.models import GIN
    from dummynode4graphlearning_b200.optim import FlatAdam
    from dummynode4graphlearning_b200.pipelines import ClassificationPipeline, _CapturedTransform, pin_batch
    raws = [{k: v for k, v in synth.tu_batch("proteins", 48, seed=s).items() if k != "vattr"} for s in (1, 2)]
    order = [0, 0, 0, 1, 1, 0, 1, 0, 0, 1]
    args = Namespace(num_features=2, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                     additional={"train_eps": True, "num_layers": 3, "aggregation": "sum"}, epochs=1, device=str(device))

    def run(hints, host_api):
        torch.manual_seed(0)
        model = GIN(args).to(device)
        pipe = ClassificationPipeline(model, FlatAdam(model.parameters(), lr=0.003), mode="conj", num_node_labels=2,
                                      node_label_min=0, cuda_graphs=True)
        bs = [dict(r, conj_sizes=T.tu_conjugate_sizes(r, True)) if hints else r for r in raws]
        if host_api:
            hosts = [pin_batch(b) for b in bs]
            losses = [pipe.step(hosts[i]) for i in order]
        else:
            devs = [T.to_device(b, device) for b in bs]
            losses = [float(pipe.step_resident(devs[i]).item()) for i in order]
        torch.cuda.synchronize()
        return losses, pipe

    ref, _ = run(False, False)
    for host_api in (False, True):
        got, pipe = run(True, host_api)
        caps = [e for e in pipe._tgraphs.values() if isinstance(e, _CapturedTransform)]
        assert len(caps) == 2 and all(c.replays >= 2 for c in caps), "the transform was not replayed from a CUDA graph"
        for a, b in zip(ref, got):
            assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (host_api, ref, got)
    # structure tensors of a replay == eager transform of the same batch
    _, pipe = run(True, False)
    dev = T.to_device(dict(raws[0], conj_sizes=T.tu_conjugate_sizes(raws[0], True)), device)
    cap = pipe.transform(dev)
    eager = pipe._transform_eager(dev)
    torch.cuda.synchronize()
    for name in ("csr_in", "csr_out"):
        a, b = getattr(cap.structure, name), getattr(eager.structure, name)
        n = b.n_rows
        assert torch.equal(a.row_ptr[: n + 1], b.row_ptr[: n + 1])
        nnz = int(b.row_ptr[n])
        assert torch.equal(a.col[:nnz], b.col[:nnz])
    assert torch.equal(cap.x, eager.x) and torch.equal(cap.y, eager.y)
