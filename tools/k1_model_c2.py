"""CPU model of the aggregation kernel's consumer loop on the C2 batch (no GPU): how much of a warp's work is lost to
sub-groups of one warp walking rows of different length (D = 32: 8 lanes per row, 4 rows per warp iteration), and what a
degree-sorted row order inside every tile would recover.  Cost model per row: ceil(deg / 4) four-neighbour batches + the
remainder one by one (process_tile_fast, csrc/spmm_tiled.cu); rows above 64 neighbours are split over the 4 sub-groups.
Usage: python tools/k1_model_c2.py [window_rows]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dummynode4graphlearning_b200 import synth  # noqa: E402
from oracle import transforms as OT  # noqa: E402  (dev tool: the oracle builds the C2 structure on the host)


def row_cost(deg):
    deg = np.asarray(deg)
    return deg // 4 + deg % 4 + 1          # +1: row_ptr loads / store


def main():
    window = int(sys.argv[1]) if len(sys.argv) > 1 else 311
    raw = synth.tu_batch("proteins", 1113, seed=0)
    conj = OT.tu_conjugate(OT.tu_add_dummy(raw))
    s, d, _, _ = OT.pyg_coalesce(conj["src"], conj["dst"])
    N = int(conj["node_ptr"][-1])
    deg = np.bincount(d, minlength=N)
    print("rows %d, nnz %d, mean degree %.2f, p50 %d, p90 %d, p99 %d, max %d, rows > 64: %d (%.1f %% of nnz)"
          % (N, len(s), deg.mean(), np.percentile(deg, 50), np.percentile(deg, 90), np.percentile(deg, 99), deg.max(),
             (deg > 64).sum(), 100.0 * deg[deg > 64].sum() / len(s)))
    node_ptr = conj["node_ptr"].astype(np.int64)
    # graph-aligned tiles as dn4gl_make_row_tiles cuts them (boundary = first graph start inside the window, else the cut)
    T = (N + window - 1) // window
    bounds = []
    for k in range(T):
        g = np.searchsorted(node_ptr, k * window, side="left")
        gs = node_ptr[g] if g < len(node_ptr) else N
        bounds.append(int(gs) if gs < (k + 1) * window else k * window)
    bounds.append(N)
    NCW, RPW = 31, 4
    tot_mean = tot_max = tot_sorted = 0.0
    per_tile = []
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        dd = deg[r0:r1]
        light = np.where(dd > 64, 0, dd)                 # long rows: split over the sub-groups afterwards
        heavy = dd[dd > 64]
        for order, acc in ((light, "plain"), (np.sort(light)[::-1], "sorted")):
            c = row_cost(order)
            pad = (-len(c)) % RPW
            c4 = np.concatenate([c, np.zeros(pad, c.dtype)]).reshape(-1, RPW)
            if acc == "plain":
                t_mean, t_max = c4.sum() / RPW, c4.max(axis=1).sum()
            else:
                t_sorted = c4.max(axis=1).sum()
        hv = (np.ceil(heavy / (4.0 * RPW)) + 3).sum()   # per long row: every sub-group walks deg / 4 of it, 4 in flight
        tot_mean += t_mean + hv
        tot_max += t_max + hv
        tot_sorted += t_sorted + hv
        # warp-level imbalance inside the tile: iterations are dealt round-robin to NCW warps
        per_tile.append((t_max + hv) / NCW)
    print("tiles %d (window %d rows); warp-iterations of work per tile, mean %.1f" % (len(per_tile), window, np.mean(per_tile)))
    print("divergence: cost with rows as stored %.0f vs ideal (no idle sub-groups) %.0f -> %.2fx; degree-sorted inside "
          "the tile %.0f -> %.2fx of ideal" % (tot_max, tot_mean, tot_max / tot_mean, tot_sorted, tot_sorted / tot_mean))


if __name__ == "__main__":
    main()
