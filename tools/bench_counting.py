"""C3 / C4 train-step throughput of the counting pipelines (BASELINE.json configs[2], configs[3]); not the contract
bench line (bench.py reports C2) -- numbers for DESIGN.md.

  python tools/bench_counting.py --config c3|c4 [--steps 20] [--no-graphs] [--cpu]
Multi-GPU: launch with torch.distributed.run like bench.py (weak scaling: every rank owns its own batch)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

CFG = {"c3": ("RGIN", "small", 512, dict(rep_rgin_regularizer="bdd", rep_rgin_num_bases=4)),
       "c4": ("DMPNN", "large", 64, dict(node_pred=True, edge_pred=False))}


def build(name, shape, over, device):
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.subgraph_isomorphism.models import DMPNN, RGIN
    cfg = dict(synth.counting_config(shape), add_dummy=True)
    mc = T.process_model_config(cfg)
    kw = dict({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=64, rep_num_graph_layers=3,
              rep_num_pattern_layers=3, rep_act_func="relu", pred_act_func="relu", pred_net="SumPredictNet",
              pred_hid_dim=64, emb_net="Equivariant", enc_net="Multihot", filter_net="ScalarFilter", pred_with_enc=True,
              pred_with_deg=True, init_neigenv=4.0, init_eeigenv=4.0)
    kw.update(over)
    torch.manual_seed(0)
    model = {"RGIN": RGIN, "DMPNN": DMPNN}[name](**kw)
    with torch.no_grad():
        for n, q in model.named_parameters():
            if "pred_fc2" in n or "weight_fc2" in n:
                q.normal_(0.0, 0.1)
    return model.to(device), cfg, kw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--opt", default="flat", choices=["flat", "torch"], help="FlatAdam (dn4gl_adam_f32) or torch.optim.AdamW")
    ap.add_argument("--no-overlap", action="store_true")
    a = ap.parse_args()
    import torch.distributed as dist
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.parallel import max_over_ranks
    from dummynode4graphlearning_b200.pipelines import CountingPipeline
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name, shape, bs, over = CFG[a.config]
    model, cfg, kw = build(name, shape, over, dev)
    if a.opt == "flat":     # train.py:1408-1411 (AdamW, amsgrad) as one flat-buffer kernel
        from dummynode4graphlearning_b200.optim import FlatAdam
        opt = FlatAdam(model.parameters(), lr=1e-3, weight_decay=1e-2, amsgrad=True, decoupled_weight_decay=True)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, amsgrad=True, capturable=True)
    pipe = CountingPipeline(model, opt, cfg, add_dummy=True, rep_reg_w=1e-3, cuda_graphs=not a.no_graphs,
                            overlap=False if a.no_overlap else None)
    pipe.global_batch = bs * world
    p, g, counts = synth.counting_batch(shape, bs, seed=rank)
    pd_, gd_ = T.to_device(p, dev), T.to_device(g, dev)
    cd = torch.from_numpy(counts).to(dev)
    for _ in range(a.warmup):
        loss = pipe.step_resident(pd_, gd_, cd, assume_ready=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = pipe.step_resident(pd_, gd_, cd, assume_ready=True)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / a.steps
    if rank == 0:
        print(json.dumps({"config": a.config, "model": name, "shape": shape, "graphs_per_gpu": bs, "n_gpus": world,
                          "cuda_graphs": not a.no_graphs, "optimizer": a.opt, "overlap": pipe.overlap, "ms_per_step": ms, "graphs_per_s": bs * world / (ms * 1e-3),
                          "loss": float(loss.item()), "replayed_library_kernels": pipe.replayed_library_kernels()}), flush=True)
    if world > 1:
        pipe._graphs.clear()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
