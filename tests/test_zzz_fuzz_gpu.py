"""GPU: the graph-construction kernels against the oracle on inputs the synthetic shapes never produce -- self loops,
duplicate (multi-) edges, isolated and single nodes, graphs without edges, one-graph batches -- bit-exact.  The oracle
is pinned to the unmodified reference on the same generators in tests/test_oracle_vs_reference.py::*_fuzz_live."""
import numpy as np
import pytest
import torch

from helpers import NASTY_CFG, nasty_sub_batch, nasty_tu_batch, oracle_preprocess_chain
from oracle import transforms as OT

pytestmark = pytest.mark.gpu

TU = ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "vid", "eid", "v_is_dummy", "e_is_dummy")
SUB = ("node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "eid", "elabel", "v_is_dummy", "e_is_dummy", "e_is_reversed",
       "v_is_reversed")


def _same(mine, ref, keys, what):
    for k in keys:
        if k in ref:
            assert k in mine, (what, k)
            got = mine[k].cpu().numpy() if isinstance(mine[k], torch.Tensor) else np.asarray(mine[k])
            assert np.array_equal(got.astype(np.int64), np.asarray(ref[k]).astype(np.int64)), (what, k)


def _no_async_errors():
    from dummynode4graphlearning_b200.graph import check_errors
    torch.cuda.synchronize()
    check_errors()            # capacity / ordering violations the builder kernels report through the device flag


@pytest.mark.parametrize("seed", range(16))
def test_tu_transforms_fuzz(device, seed):
    from dummynode4graphlearning_b200 import transforms as T
    rng = np.random.default_rng(seed)
    b = nasty_tu_batch(rng, int(rng.integers(1, 6)))
    dev = T.to_device(b, device)
    dummy = T.tu_add_dummy(dev)
    _same(dummy, OT.tu_add_dummy(b), TU, "dummy")
    _same(T.tu_conjugate(dummy), OT.tu_conjugate(OT.tu_add_dummy(b)), TU, "conj")
    _same(T.tu_conjugate(dev), OT.tu_conjugate(b), TU, "line")
    # canonical (PyG) view of the CONJ graphs: remove_self_loops + coalesce
    oc = OT.tu_conjugate(OT.tu_add_dummy(b))
    es, ed, _, _ = OT.pyg_coalesce(oc["src"], oc["dst"])
    can = T.pyg_canonicalize(T.tu_conjugate(T.tu_add_dummy(T.to_device(b, device))), with_edge_attr=False)
    assert np.array_equal(can["edge_index"].cpu().numpy(), np.stack([es, ed]).astype(np.int64))
    _no_async_errors()


@pytest.mark.parametrize("seed", range(16))
def test_sub_preprocessing_fuzz(device, seed):
    from dummynode4graphlearning_b200 import transforms as T
    from dummynode4graphlearning_b200.pipelines import CountingPipeline
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 5))
    p, g = nasty_sub_batch(rng, B, 5, 4), nasty_sub_batch(rng, B, 8, 6)
    flags = tuple(bool(x) for x in rng.integers(0, 2, 4))
    lin = torch.nn.Linear(2, 2).to(device)
    pipe = CountingPipeline(lin, torch.optim.SGD(lin.parameters(), lr=0.0), NASTY_CFG, add_dummy=flags[2], cuda_graphs=False,
                            remove_loops=flags[0], add_rev=flags[1], convert_conj=flags[3], share_emb_net=False)
    mp, mg = pipe.augment(T.to_device(p, device), T.to_device(g, device))
    op, og = oracle_preprocess_chain(p, g, NASTY_CFG, *flags)
    _same(mp, op, SUB, ("pattern", flags))
    _same(mg, og, SUB, ("graph", flags))
    _no_async_errors()


def test_match_weights_fuzz(device):
    """batched match-weight / conjugate-map kernels on multigraphs with self loops and repeated (u, v, label) edges."""
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.subgraph_isomorphism import matching as M
    done = 0
    for seed in range(3000, 3200):
        rng = np.random.default_rng(seed)
        B = int(rng.integers(1, 5))
        p, g = nasty_sub_batch(rng, B, 5, 4), nasty_sub_batch(rng, B, 8, 6)
        if (np.diff(p["edge_ptr"]) == 0).any() or (np.diff(g["edge_ptr"]) == 0).any():
            continue
        mats = synth.random_subisomorphisms(p, g, seed=seed, max_rows=6)
        sub = M.pack_subisomorphisms(mats, device)
        pb, gb = T.to_device(p, device), T.to_device(g, device)
        assert np.array_equal(M.node_weights(sub, gb).cpu().numpy(), OT.subiso_node_weights(mats, g)), seed
        assert np.array_equal(M.edge_weights(sub, pb, gb).cpu().numpy(), OT.subiso_edge_weights(mats, p, g)), seed
        work, conj = M.conjugate_subisomorphisms(sub, pb, gb)
        work, conj = work.cpu().numpy(), conj.cpu().numpy()
        for b, ref in enumerate(OT.subiso_conjugate(mats, p, g)):
            assert np.array_equal(conj[work[b]: work[b + 1]].reshape(ref.shape), ref), (seed, b)
        done += 1
        if done == 15:
            break
    assert done == 15
    _no_async_errors()
