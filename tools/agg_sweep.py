#!/usr/bin/env python
"""C5 aggregation sweep (BASELINE.json configs[4]): batches of 1k-64k MUTAG-shaped graphs + dummy node,
hidden 64-512, forward sum aggregation through the C ABI, device-timed with an L2 flush before every launch.

  python tools/agg_sweep.py [--graphs 1024,4096,...] [--dims 64,128,...] [--modes rows,tiled] [--out file.json]

Prints one JSON object per (B, D, mode): algorithmic GB/s (4DN*2 + 4E + 4(N+1) bytes) and its
fraction of the measured HBM copy peak.  Large batches are made by block-diagonal replication of a 1024-graph
seeded base batch (graphs are independent, so the replica is a valid mini-batch of the named shape).
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def replicate(b, times):
    """block-diagonal concatenation of `times` copies of a flat batch dict (numpy)."""
    if times == 1:
        return b
    N, E = int(b["node_ptr"][-1]), int(b["edge_ptr"][-1])
    out = dict(b)
    out["num_graphs"] = b["num_graphs"] * times
    out["node_ptr"] = np.concatenate([[0]] + [b["node_ptr"][1:].astype(np.int64) + k * N for k in range(times)]).astype(np.int32)
    out["edge_ptr"] = np.concatenate([[0]] + [b["edge_ptr"][1:].astype(np.int64) + k * E for k in range(times)]).astype(np.int32)
    for key in ("src", "dst"):
        out[key] = np.concatenate([b[key].astype(np.int64) + k * N for k in range(times)]).astype(np.int32)
    for key in ("vlabel", "elabel", "y", "vattr", "vid", "eid"):
        if key in b:
            out[key] = np.tile(b[key], times)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", default="1024,4096,16384,65536")
    ap.add_argument("--dims", default="64,128,256,512")
    ap.add_argument("--modes", default="rows,tiled")
    ap.add_argument("--smem", default="200")
    ap.add_argument("--shape", default="mutag")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--warps", default="32", help="warps per CTA of the tiled kernel (16,24,32)")
    ap.add_argument("--known", type=int, default=1, help="1: the host knows the largest graph (graph-aligned tiles)")
    a = ap.parse_args()

    from dummynode4graphlearning_b200 import graph as graph_mod, ops, synth, transforms as T
    from dummynode4graphlearning_b200.graph import BatchedGraph

    dev = torch.device("cuda:0")
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = synth.tu_batch(a.shape, 1024, seed=0)
    results = []
    for B in [int(x) for x in a.graphs.split(",")]:
        raw = replicate(base, max(B // 1024, 1)) if B >= 1024 else synth.tu_batch(a.shape, B, seed=0)
        d = T.tu_add_dummy(T.to_device({k: v for k, v in raw.items() if k != "vattr"}, dev))
        g = BatchedGraph(d["src"], d["dst"], d["node_ptr"], d["edge_ptr"])
        N, E = g.number_of_nodes(), g.number_of_edges()
        if a.known:
            g.host_ptrs()
        g.csr_in, g.csr_out
        for D in [int(x) for x in a.dims.split(",")]:
            if 4 * D * N * 3 > 60e9:
                continue
            x = torch.rand((N, D), device=dev) * 2 - 1
            bytes_alg = 4 * D * N * 2 + 4 * E + 4 * (N + 1)
            for mode in a.modes.split(","):
                cfgs = [(int(s), int(w)) for s in a.smem.split(",") for w in a.warps.split(",")] if mode == "tiled" else [(0, 0)]
                for smem_kb, warps in cfgs:
                    ops.SPMM_MODE = mode
                    if smem_kb:
                        ops.TILE_SMEM = smem_kb * 1024
                        graph_mod.TILE_WARPS = warps
                        g.csr_in._tiles.clear()
                    for _ in range(3):
                        ops.graph_sum_aggregate(g, x, 1.0)
                    ts = []
                    for _ in range(a.iters):
                        flush.fill_(1)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        ops.graph_sum_aggregate(g, x, 1.0)
                        e1.record()
                        torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    us = 1e3 * statistics.median(ts)
                    r = {"graphs": B, "N": N, "E": E, "D": D, "mode": mode, "smem_kb": smem_kb, "warps": warps, "us": round(us, 2),
                         "alg_MB": round(bytes_alg / 1e6, 2), "alg_GBs": round(bytes_alg / us / 1e3, 1),
                         "frac_of_measured_peak": round(bytes_alg / us / 1e3 / peak, 4),
                         "graphs_per_s": round(B / us * 1e6)}
                    results.append(r)
                    print(json.dumps(r), flush=True)
            del x
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"peak_hbm_gbs": peak, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
