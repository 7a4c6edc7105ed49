#!/usr/bin/env python
"""Per-role wait accounting of the pipelined MLP stage kernels (debug build: make libdn4gl_exp.so EXP_FLAGS=-DDN4GL_PIPE_TL,
DN4GL_LIB=.../libdn4gl_pipetl.so python tools/pipe_timeline.py [--rows N] [--dims 32,64] [--op fwd|fwd_bn]).
For every role (producer, MMA issuer, converters, epilogue): cycles of its tile loop and the share it spent blocked in each
of its waits, averaged over the CTAs.  The role that waits least bounds the pipeline."""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dummynode4graphlearning_b200 import ops, _lib

ROLES = {"bwd": [("producer", ["raw_empty"]), ("mma", ["acc_empty", "a_full"]), ("convert", ["raw_full", "a_empty"]),
                 ("epilogue", ["acc_full", "PHASE_work"])],
         "fwd": [("producer", ["raw_empty"]), ("mma", ["acc_empty", "a_full"]), ("convert", ["raw_full", "a_empty"]),
                 ("epilogue", ["acc_full", "PHASE_tmem_to_staging", "PHASE_store"])]}


def read():
    buf = (ctypes.c_longlong * (148 * 16))()
    rc = _lib.lib().raw("dn4gl_debug_read_pipe_timeline")(buf)
    assert rc == 0
    return torch.tensor(list(buf), dtype=torch.float64).view(148, 4, 4)


def read_span():
    buf = (ctypes.c_ulonglong * (148 * 4))()
    assert _lib.lib().raw("dn4gl_debug_read_pipe_span")(buf) == 0
    return torch.tensor([float(x) for x in buf], dtype=torch.float64).view(148, 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="156759,1000000")
    ap.add_argument("--dims", default="32,64")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    for N in [int(x) for x in a.rows.split(",")]:
        for D in [int(x) for x in a.dims.split(",")]:
            X = torch.randn(N, D, device=dev)
            W = torch.randn(D, D, device=dev) / D ** 0.5
            b = torch.randn(D, device=dev)
            bn = dict(gamma=torch.ones(D, device=dev), beta=torch.zeros(D, device=dev), eps=1e-5, momentum=0.1)
            G = torch.randn(N, D, device=dev)
            Y, rec = ops.lin_fwd(X, W, b, bn=bn)
            sums = ops.bn_bwd_sums(G, Y, rec)
            for op in ("fwd", "fwd_bn", "bwd", "bwd_bn"):
                for _ in range(3):
                    if op.startswith("fwd"):
                        ops.lin_fwd(X, W, b, bn=bn if op == "fwd_bn" else None)
                    elif op == "bwd":
                        ops.lin_bwd(G, W, X)
                    else:
                        ops.lin_bwd(G, W, X, Yout=Y, bn=rec, sums=sums, in_bn=rec, in_act=1)
                torch.cuda.synchronize()
                tl = read()
                ctas = min(148, (N + 127) // 128)
                out = {"N": N, "D": D, "op": op, "tiles_per_cta": round((N + 127) // 128 / ctas, 2)}
                for r, (name, waits) in enumerate(ROLES["fwd" if op.startswith("fwd") else "bwd"]):
                    tot = tl[:ctas, r, 0]
                    out[name] = {"loop_cycles_mean": round(float(tot.mean())), "loop_cycles_max": round(float(tot.max()))}
                    for k, wn in enumerate(waits):
                        out[name]["wait_" + wn] = round(float((tl[:ctas, r, 1 + k] / tot.clamp_min(1)).mean()), 3)
                sp = read_span()[:ctas]
                t0 = float(sp[:, 0].min())
                us = lambda x: round((float(x) - t0) / 1e3, 2)
                out["span_us"] = {"entry_max": us(sp[:, 0].max()), "setup_done_mean": us(sp[:, 1].mean()), "setup_done_max": us(sp[:, 1].max()),
                                  "roles_done_mean": us(sp[:, 2].mean()), "roles_done_max": us(sp[:, 2].max()),
                                  "exit_max": us(sp[:, 3].max())}
                print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
