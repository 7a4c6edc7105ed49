#!/usr/bin/env python
"""Per-role wait accounting of dn4gl_gemm_f32 (debug build: make -C dummynode4graphlearning_b200/csrc libdn4gl_exp.so
EXP_FLAGS=-DDN4GL_GEMM_TL; DN4GL_LIB=.../libdn4gl_exp.so python tools/gemm_timeline.py): cycles of each role's chunk loop and
the share blocked in each wait, averaged over the CTAs."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dummynode4graphlearning_b200 import ops, _lib

ROLES = [("producer", ["a_empty"]), ("mma", ["acc_empty", "a_full"]), ("convert", ["raw_full", "a_empty", "FENCE"]),
         ("loader", ["raw_empty"]), ("epilogue", ["acc_full", "WORK_drain", "WORK_store"])]


def main():
    dev = torch.device("cuda:0")
    for N, K, M in [(1000000, 64, 64), (156759, 256, 256), (15424, 1024, 64)]:
        a, b = torch.randn(N, K, device=dev), torch.randn(M, K, device=dev)
        for _ in range(3):
            ops.gemm(a, b, 0)
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (148 * 5 * 4))()
        assert _lib.lib().raw("dn4gl_debug_read_gemm_timeline")(buf) == 0
        tl = torch.tensor(list(buf), dtype=torch.float64).view(148, 5, 4)
        MT = 16 if M <= 16 else 32 if M <= 32 else 64 if M <= 64 else 128
        tiles = ((N + 127) // 128) * ((M + MT - 1) // MT)
        ctas = min(148, tiles)
        chunks = tiles / ctas * ((K + 31) // 32)
        out = ["N=%d K=%d M=%d: %.1f chunks per CTA" % (N, K, M, chunks)]
        for r, (name, waits) in enumerate(ROLES):
            tot = tl[:ctas, r, 0]
            line = "  %-9s loop %7.0f cycles (%5.0f per chunk)" % (name, float(tot.mean()), float(tot.mean()) / chunks)
            for k, wn in enumerate(waits):
                line += "  %s %.2f" % (wn, float((tl[:ctas, r, 1 + k] / tot.clamp_min(1)).mean()))
            out.append(line)
        print("\n".join(out), flush=True)


if __name__ == "__main__":
    main()
