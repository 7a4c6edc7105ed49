#!/usr/bin/env python
"""dn4gl_gemm_f32 against the library fp32 GEMM on shapes of the C3 / C4 models: time per call from one CUDA-graph replay of
16 back-to-back calls (the eager loop would time the Python -> ctypes call at these sizes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dummynode4graphlearning_b200 import ops

SHAPES = [(15424, 64, 64, 0), (15424, 64, 1024, 1), (15424, 1024, 64, 0), (15424, 64, 256, 1), (38372, 64, 64, 0),
          (45397, 64, 128, 1), (45397, 128, 64, 0), (17472, 64, 128, 1), (17472, 256, 64, 0), (512, 64, 64, 0), (512, 256, 1, 0),
          (156759, 128, 128, 0), (156759, 256, 256, 0), (1000000, 64, 64, 0)]


def graph_time(fn, reps=16):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (5 * reps)


def main():
    dev = torch.device("cuda:0")
    for N, K, M, layout in SHAPES:
        a = torch.randn(N, K, device=dev)
        b = torch.randn(M, K, device=dev) if layout == 0 else torch.randn(K, M, device=dev)
        t_mine = graph_time(lambda: ops.gemm(a, b, layout))
        t_lib = graph_time(lambda: (a @ b.t()) if layout == 0 else (a @ b))
        fl = 2.0 * N * K * M
        print("N=%7d K=%4d M=%4d layout %d: dn4gl %7.2f us (%6.1f TFLOP/s)   library fp32 %7.2f us (%6.1f TFLOP/s)"
              % (N, K, M, layout, t_mine, fl / t_mine * 1e-6, t_lib, fl / t_lib * 1e-6), flush=True)


if __name__ == "__main__":
    main()
