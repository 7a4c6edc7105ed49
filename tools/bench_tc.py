"""micro-benchmark of the tensor-core MLP stages: device time per launch (CUDA events around a loop, L2 flushed
between launches by cycling through distinct buffers), algorithmic GB/s."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dummynode4graphlearning_b200 import ops


def timeit(fn, nbuf, iters=20):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i % nbuf)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="156759,1000000")
    ap.add_argument("--dims", default="32,64")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    res = []
    for N in [int(x) for x in a.rows.split(",")]:
        for D in [int(x) for x in a.dims.split(",")]:
            nbuf = max(2, int(400e6 // (N * D * 4)) + 1)   # > L2 worth of distinct inputs
            nbuf = min(nbuf, 12)
            X = [torch.randn(N, D, device=dev) for _ in range(nbuf)]
            G = [torch.randn(N, D, device=dev) for _ in range(nbuf)]
            W = torch.randn(D, D, device=dev) / D ** 0.5
            b = torch.randn(D, device=dev)
            gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
            bn = dict(gamma=gam, beta=bet, eps=1e-5, momentum=0.1)
            Y, rec = ops.lin_fwd(X[0], W, b, bn=bn)
            sums = ops.bn_bwd_sums(G[0], Y, rec)
            rows = {}
            rows["fwd_plain"] = (timeit(lambda i: ops.lin_fwd(X[i], W, b), nbuf), 2 * N * D * 4)
            rows["fwd_bn_stats"] = (timeit(lambda i: ops.lin_fwd(X[i], W, b, in_bn=rec, in_act=1, bn=bn), nbuf), 2 * N * D * 4)
            rows["bwd_plain"] = (timeit(lambda i: ops.lin_bwd(G[i], W, X[i]), nbuf), 3 * N * D * 4)
            rows["bwd_bn"] = (timeit(lambda i: ops.lin_bwd(G[i], W, X[i], Yout=Y, bn=rec, sums=sums, in_bn=rec, in_act=1), nbuf), 4 * N * D * 4)
            rows["bwd_nogx"] = (timeit(lambda i: ops.lin_bwd(G[i], W, X[i], want_gx=False), nbuf), 2 * N * D * 4)
            rows["bn_act"] = (timeit(lambda i: ops.bn_act(X[i], rec), nbuf), 2 * N * D * 4)
            rows["bn_bwd_sums"] = (timeit(lambda i: ops.bn_bwd_sums(G[i], X[i], rec), nbuf), 2 * N * D * 4)
            rows["torch_addmm"] = (timeit(lambda i: torch.addmm(b, X[i], W.t()), nbuf), 2 * N * D * 4)
            for k, (us, byts) in rows.items():
                r = dict(N=N, D=D, op=k, us=round(us, 2), alg_gbs=round(byts / us / 1e3, 1))
                print(json.dumps(r), flush=True)
                res.append(r)
            del X, G
            torch.cuda.empty_cache()
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
