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
