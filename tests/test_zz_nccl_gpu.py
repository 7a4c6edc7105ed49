"""GPU, two ranks over NCCL (skipped on a one-GPU box; run with ``gpurun --gpus 2``): a mini-batch sharded over two
processes -- one per GPU, each building its own CSR, gradients all-reduced through ONE flat bucket with weights
B_r / B by the peer-memory kernel of csrc/peer.cu (pipelines.py / parallel.py, the path bench.py takes at N > 1) -- against the SAME mini-batch evaluated by a
single process on one GPU:

* C3-style counting step (RGIN + dummy, no BatchNorm, ``exact_sharding=True``: padded lengths agreed over NCCL):
  the all-reduced gradient equals the single-process gradient of the whole batch within 1e-5;
* C2-style classification step (GIN with BatchNorm): BatchNorm statistics are per rank (documented in DESIGN.md
  section 7), so the reference value is the B_r / B weighted sum of the two shards' single-process gradients -- which
  pins the bucket weighting and the collective, not a cross-rank BatchNorm.
"""
import os
import socket
import tempfile
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _slice_tu(b, lo, hi):
    n0, n1, e0, e1 = int(b["node_ptr"][lo]), int(b["node_ptr"][hi]), int(b["edge_ptr"][lo]), int(b["edge_ptr"][hi])
    out = dict(num_graphs=hi - lo, node_ptr=(b["node_ptr"][lo:hi + 1] - n0).astype(np.int32),
               edge_ptr=(b["edge_ptr"][lo:hi + 1] - e0).astype(np.int32),
               src=(b["src"][e0:e1] - n0).astype(np.int32), dst=(b["dst"][e0:e1] - n0).astype(np.int32))
    for k in ("vlabel", "vid", "v_is_dummy"):
        if k in b:
            out[k] = b[k][n0:n1]
    for k in ("elabel", "eid", "e_is_dummy", "e_is_reversed"):
        if k in b:
            out[k] = b[k][e0:e1]
    if "y" in b:
        out["y"] = b["y"][lo:hi]
    return out


def _counting_setup(device):
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.subgraph_isomorphism.models import RGIN
    cfg = dict(synth.counting_config("small"), add_dummy=True)
    mc = T.process_model_config(cfg)
    kw = dict({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=64, rep_num_graph_layers=2,
              rep_num_pattern_layers=2, rep_act_func="relu", pred_act_func="relu", pred_net="SumPredictNet",
              pred_hid_dim=32, emb_net="Equivariant", enc_net="Multihot", filter_net="ScalarFilter", pred_with_enc=True,
              pred_with_deg=True, rep_rgin_regularizer="bdd", rep_rgin_num_bases=4)
    torch.manual_seed(3)
    model = RGIN(**kw)
    with torch.no_grad():
        for n, q in model.named_parameters():
            if "pred_fc2" in n or "weight_fc2" in n:
                q.normal_(0.0, 0.1)
    return model.to(device), cfg


def _counting_grads(device, lo, hi, exact, total):
    """one eager step (lr 0) of samples [lo, hi) of the seeded batch -> flat gradient after the (possible) all-reduce"""
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.pipelines import CountingPipeline
    model, cfg = _counting_setup(device)
    p, g, counts = synth.counting_batch("small", total, seed=11)
    pipe = CountingPipeline(model, torch.optim.SGD(model.parameters(), lr=0.0), cfg, add_dummy=True, rep_reg_w=1e-3,
                            max_grad_norm=0.0, cuda_graphs=False, exact_sharding=exact)
    pipe.global_batch = total
    ps, gs = _slice_tu(p, lo, hi), _slice_tu(g, lo, hi)
    loss = pipe.step_resident(T.to_device(ps, device), T.to_device(gs, device), torch.from_numpy(counts[lo:hi]).to(device))
    return float(loss), pipe.bucket.flat.detach().cpu().clone()


def _classification_grads(device, lo, hi, total):
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.graph_classification.models import GIN
    from dummynode4graphlearning_b200.pipelines import ClassificationPipeline
    raw = synth.tu_batch("proteins", total, seed=4)
    raw = {k: v for k, v in raw.items() if k != "vattr"}
    args = Namespace(num_features=2, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                     additional={"train_eps": True, "num_layers": 3, "aggregation": "sum"}, epochs=1, device=str(device))
    torch.manual_seed(5)
    model = GIN(args).to(device)
    pipe = ClassificationPipeline(model, torch.optim.SGD(model.parameters(), lr=0.0), mode="conj", num_node_labels=2,
                                  node_label_min=0, cuda_graphs=False, overlap=False)
    pipe.global_batch = total
    loss = pipe.step_resident(T.to_device(_slice_tu(raw, lo, hi), device))
    return float(loss), pipe.bucket.flat.detach().cpu().clone()


def _worker(rank, world, port, out_dir, total_c, total_g):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        half_c, half_g = total_c // 2, total_g // 2
        lc, gc = _counting_grads(dev, rank * half_c, (rank + 1) * half_c, True, total_c)
        lg, gg = _classification_grads(dev, rank * half_g, (rank + 1) * half_g, total_g)
        res = dict(loss_c=lc, grad_c=gc, loss_g=lg, grad_g=gg)
        # the peer-memory all-reduce on its own: small (one block) and large (32 blocks) buckets, three rounds each so the
        # epochs advance and the buckets are rewritten in between; against the library collective on the same data
        from dummynode4graphlearning_b200.parallel import PeerAllReduce
        for tag, n in (("small", 4 * 1037), ("large", 4 * 75001)):
            torch.manual_seed(100 + rank)
            ref = torch.randn(n, device=dev)
            flat = torch.empty_like(ref)
            peer = PeerAllReduce.create(flat)
            res["peer_" + tag] = bool(peer)
            if not peer:
                continue
            w = 0.25 + 0.5 * rank
            errs = []
            for it in range(3):
                flat.copy_(ref * (it + 1))
                peer.run(w)
                expect = ref * (it + 1) * w
                dist.all_reduce(expect)
                errs.append(float((flat - expect).abs().max() / expect.abs().max()))
            res["peer_err_" + tag] = max(errs)
            res["peer_out_" + tag] = flat.cpu()
        torch.save(res, os.path.join(out_dir, "rank%d.pt" % rank))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_step_equals_single_process(device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from helpers import rel_err
    total_c, total_g = 32, 24
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d, total_c, total_g), nprocs=2, join=True)
        r0, r1 = torch.load(os.path.join(d, "rank0.pt")), torch.load(os.path.join(d, "rank1.pt"))
    # both ranks hold the same reduced bucket
    assert torch.equal(r0["grad_c"], r1["grad_c"]) and torch.equal(r0["grad_g"], r1["grad_g"])
    # the exchange ran through dn4gl_peer_allreduce_f32 (NVLink peer memory), not through the library fallback ...
    for tag in ("small", "large"):
        assert r0["peer_" + tag] and r1["peer_" + tag], "peer-memory all-reduce was not set up"
        assert r0["peer_err_" + tag] <= 1e-6 and r1["peer_err_" + tag] <= 1e-6, (r0["peer_err_" + tag], r1["peer_err_" + tag])
        assert torch.equal(r0["peer_out_" + tag], r1["peer_out_" + tag])     # ... and every rank computed the same bits
    # counting (no BatchNorm, padded lengths agreed): == the single-process gradient of the whole batch
    _, ref_c = _counting_grads(device, 0, total_c, False, total_c)
    e = rel_err(r0["grad_c"], ref_c)
    assert e <= 1e-5, "sharded RGIN gradient vs single process: %.2e" % e
    # classification (per-rank BatchNorm statistics): == sum_r (B_r / B) * single-process gradient of shard r
    half = total_g // 2
    _, s0 = _classification_grads(device, 0, half, total_g)        # single process: unweighted shard gradients
    _, s1 = _classification_grads(device, half, total_g, total_g)
    w = half / total_g
    e = rel_err(r0["grad_g"], w * s0 + w * s1)
    assert e <= 1e-5, "sharded GIN gradient vs weighted shard gradients: %.2e" % e
