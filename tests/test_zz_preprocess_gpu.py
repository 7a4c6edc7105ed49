"""GPU: CountingPipeline.augment -- the reference's data-set level preprocessing switches chained per mini-batch on
the GPU (remove_loops -> add_reversed_edges -> add_dummy_nodes_edges -> convert_to_conjugate, train.py:1271-1340) --
bit-exact against the oracle chain, which tests/test_oracle_vs_reference.py::test_preprocessing_chain_live pins to the
reference's own functions chained the same way; then one train step on the fully preprocessed batch."""
import numpy as np
import pytest
import torch

from helpers import oracle_preprocess_chain

pytestmark = pytest.mark.gpu

KEYS = ("node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "eid", "elabel", "v_is_dummy", "e_is_dummy", "e_is_reversed",
        "v_is_reversed")


def _batch(seed, B=9):
    from dummynode4graphlearning_b200 import synth
    p, g, counts = synth.counting_batch("small", B, seed=seed)
    g = dict(g)
    g["dst"] = g["dst"].copy()
    g["dst"][::5] = g["src"][::5]                     # plant loops
    return p, g, counts


@pytest.mark.parametrize("share_emb_net", [False, True])
@pytest.mark.parametrize("flags", [(True, True, True, True), (False, True, True, False), (True, False, True, True),
                                   (False, True, False, True), (False, False, True, False)])
def test_augment_chain_matches_oracle(device, flags, share_emb_net):
    """share_emb_net (the reference's default, train.py:1276-1289): the pattern side is augmented with the GRAPH maxima --
    the 'small' config has different pattern / graph maxima (8 / 8 / 8 / 8 vs 64 / 256 / 16 / 16), so the pattern's dummy
    and reversed ids / labels differ between the two settings."""
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.pipelines import CountingPipeline
    remove_loops, add_rev, add_dummy, convert_conj = flags
    p, g, _ = _batch(sum(flags) * 31 + 5)
    cfg = synth.counting_config("small")
    lin = torch.nn.Linear(2, 2).to(device)            # augment() needs no model; the constructor wants parameters
    pipe = CountingPipeline(lin, torch.optim.SGD(lin.parameters(), lr=0.0), cfg, add_dummy=add_dummy, cuda_graphs=False,
                            remove_loops=remove_loops, add_rev=add_rev, convert_conj=convert_conj, share_emb_net=share_emb_net)
    mp, mg = pipe.augment(T.to_device(p, device), T.to_device(g, device))
    ocfg = dict(cfg, max_npv=cfg["max_ngv"], max_npvl=cfg["max_ngvl"], max_npe=cfg["max_nge"], max_npel=cfg["max_ngel"]) \
        if share_emb_net else cfg
    op, og = oracle_preprocess_chain(p, g, ocfg, *flags)
    for mine, ref, side in ((mp, op, "pattern"), (mg, og, "graph")):
        for k in KEYS:
            if k in ref:
                assert k in mine, (side, k)
                assert np.array_equal(mine[k].cpu().numpy().astype(np.int64), np.asarray(ref[k]).astype(np.int64)), (side, k)


def test_train_step_on_fully_preprocessed_batch(device):
    """reversed edges + dummy + edge-to-vertex, model built from process_model_config of the same switches: one eager
    step runs and yields a finite loss and gradients for every trainable parameter that takes part."""
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.pipelines import CountingPipeline
    from dummynode4graphlearning_b200.subgraph_isomorphism.models import RGIN
    p, g, counts = _batch(77)
    cfg = synth.counting_config("small")
    mc = T.process_model_config(dict(cfg, add_rev=True, add_dummy=True, convert_conj=True))
    kw = dict({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=32, rep_num_graph_layers=2,
              rep_num_pattern_layers=2, pred_hid_dim=32, emb_net="Equivariant", filter_net="ScalarFilter")
    torch.manual_seed(0)
    model = RGIN(**kw).to(device)
    pipe = CountingPipeline(model, torch.optim.SGD(model.parameters(), lr=1e-3), cfg, cuda_graphs=False, add_rev=True,
                            convert_conj=True, rep_reg_w=1e-3)
    loss = pipe.step_resident(T.to_device(p, device), T.to_device(g, device), torch.from_numpy(counts).to(device))
    assert np.isfinite(float(loss))
    got = [n for n, q in model.named_parameters() if q.grad is not None]
    assert len(got) > 10 and all(bool(torch.isfinite(q.grad).all()) for _, q in model.named_parameters() if q.grad is not None)
