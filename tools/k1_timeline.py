#!/usr/bin/env python
"""Where does one K1 launch spend its time?  Per-CTA %globaltimer stamps from the debug build libdn4gl_tl.so
(make -C dummynode4graphlearning_b200/csrc libdn4gl_tl.so) on the C2 structure (or a C5 point), one cold launch.
  DN4GL_LIB=dummynode4graphlearning_b200/csrc/libdn4gl_tl.so python tools/k1_timeline.py [--c5 16384 --dim 64]"""
import argparse
import ctypes
import json
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DN4GL_LIB", os.path.join(ROOT, "dummynode4graphlearning_b200", "csrc", "libdn4gl_tl.so"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c5", type=int, default=0, help="graphs of a C5 point instead of the C2 structure")
    ap.add_argument("--dim", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from dummynode4graphlearning_b200 import _lib, ops, synth, transforms as T
    dev = torch.device("cuda:0")
    L = _lib.lib()
    if a.c5:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from agg_sweep import replicate
        from dummynode4graphlearning_b200.graph import BatchedGraph
        raw = replicate(synth.tu_batch("mutag", 1024, seed=0), max(a.c5 // 1024, 1))
        d = T.tu_add_dummy(T.to_device({k: v for k, v in raw.items() if k != "vattr"}, dev))
        g = BatchedGraph(d["src"], d["dst"], d["node_ptr"], d["edge_ptr"])
        g.host_ptrs()
        csr_in, csr_out, N = g.csr_in, g.csr_out, g.number_of_nodes()
    else:
        from dummynode4graphlearning_b200.graph_classification.models import GIN
        from dummynode4graphlearning_b200.pipelines import ClassificationPipeline
        raw = synth.tu_batch("proteins", 1113, seed=0)
        args = Namespace(num_features=2, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                         additional={"train_eps": True, "num_layers": 2, "aggregation": "sum"}, epochs=1, device=str(dev))
        model = GIN(args).to(dev)
        pipe = ClassificationPipeline(model, torch.optim.Adam(model.parameters()), mode="conj", num_node_labels=2, node_label_min=0)
        s = pipe.transform(T.to_device({k: v for k, v in raw.items() if k != "vattr"}, dev)).structure
        csr_in, csr_out, N = s.csr_in, s.csr_out, s.num_nodes
    D = a.dim
    x = torch.rand((N, D), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t = csr_in.tiles(D)
    hc = int(t["heavy_count"].item()) if t["heavy_count"] is not None else 0
    print(json.dumps({"N": N, "E": int(csr_in.nnz), "D": D, "tiles": t["T"], "stages": t["stages"], "window": t["window"],
                      "cap_rows": t["cap_rows"], "warps": t["warps"], "heavy_virtual_tiles": hc}))
    buf = (ctypes.c_ulonglong * (148 * 32))()
    for rep in range(a.reps):
        ops.spmm_sum(x, csr_in, csr_out, 1.0)
        torch.cuda.synchronize()
        flush.fill_(1)
        torch.cuda.synchronize()
        assert L.raw("dn4gl_debug_clear_timeline")() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.spmm_sum(x, csr_in, csr_out, 1.0)
        e1.record()
        torch.cuda.synchronize()
        assert L.raw("dn4gl_debug_read_timeline")(buf) == 0
        tl = np.frombuffer(buf, dtype=np.uint64).reshape(148, 32).astype(np.int64)
        t0 = tl[:, 0][tl[:, 0] > 0].min()
        rel = np.where(tl > 0, tl - t0, -1) / 1e3          # us since the first CTA entered
        entry, prol = rel[:, 0], rel[:, 1] - rel[:, 0]
        cons = rel[:, 2:16]
        n_items = (cons >= 0).sum(1) // 2
        first_full = cons[:, 0] - rel[:, 1]
        ends = np.array([cons[i, 2 * n_items[i] - 1] if n_items[i] > 0 else rel[i, 1] for i in range(148)])
        proc, wait = [], []
        for i in range(148):
            prev = rel[i, 1]
            for k in range(n_items[i]):
                wait.append(cons[i, 2 * k] - prev)
                proc.append(cons[i, 2 * k + 1] - cons[i, 2 * k])
                prev = cons[i, 2 * k + 1]
        prod = rel[:, 16:]
        out = {"event_us": round(1e3 * e0.elapsed_time(e1), 1), "entry_skew_max": round(float(entry.max()), 2),
               "prologue_mean": round(float(prol.mean()), 2), "first_full_wait_mean": round(float(first_full.mean()), 2),
               "first_full_wait_max": round(float(first_full.max()), 2), "items_per_cta_mean": round(float(n_items.mean()), 2),
               "items_per_cta_max": int(n_items.max()), "proc_mean": round(float(np.mean(proc)), 2),
               "proc_p90": round(float(np.percentile(proc, 90)), 2), "proc_max": round(float(np.max(proc)), 2),
               "wait_mean_all": round(float(np.mean(wait)), 2), "wait_after_first_mean": round(float(np.mean([w for j, w in enumerate(wait)])) , 2),
               "cta_end_mean": round(float(ends.mean()), 2), "cta_end_p10": round(float(np.percentile(ends, 10)), 2),
               "cta_end_max": round(float(ends.max()), 2),
               "producer_first_issue_mean": round(float((prod[:, 0] - rel[:, 1])[prod[:, 0] >= 0].mean()), 2)}
        print(json.dumps(out))
    # one CTA in detail: the one that finished last
    i = int(np.argmax(ends))
    print("slowest CTA", i, "stamps(us):", " ".join("%.1f" % v for v in rel[i] if v >= 0))
    i = int(np.argmin(ends))
    print("fastest CTA", i, "stamps(us):", " ".join("%.1f" % v for v in rel[i] if v >= 0))


if __name__ == "__main__":
    main()
