#!/usr/bin/env python
"""Where a C3 / C4 step goes: host time and device time of the transform (eager, second stream) and of the captured train
step, each alone, plus the step as bench.py runs it.  python tools/diag_counting.py [c3|c4]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from dummynode4graphlearning_b200 import synth, transforms as T, _lib
from dummynode4graphlearning_b200.optim import FlatAdam
from dummynode4graphlearning_b200.pipelines import CountingPipeline


GC_STATS = {"ms": 0.0, "runs": [0, 0, 0]}


def _gc_cb(phase, info):
    if phase == "start":
        GC_STATS["t0"] = time.perf_counter()
    else:
        GC_STATS["ms"] += 1e3 * (time.perf_counter() - GC_STATS["t0"])
        GC_STATS["runs"][info["generation"]] += 1


def main():
    import gc
    gc.callbacks.append(_gc_cb)
    if os.environ.get("DIAG_GC_FREEZE"):
        gc.collect(); gc.freeze()
    dev = torch.device("cuda:0")
    L = _lib.lib()
    for key in (sys.argv[1:] or ["c3", "c4"]):
        name, shape, bs, over = Bn.COUNTING[key]
        cfg, kw = Bn.counting_kwargs(shape, over)
        model = Bn.build_counting_model(name, kw, dev)
        opt = FlatAdam(model.parameters(), lr=1e-3, weight_decay=1e-2, amsgrad=True, decoupled_weight_decay=True)
        pipe = CountingPipeline(model, opt, cfg, add_dummy=True, rep_reg_w=1e-3)
        p, g, counts = synth.counting_batch(shape, bs, seed=0)
        pd_, gd_, cd = T.to_device(p, dev), T.to_device(g, dev), torch.from_numpy(counts).to(dev)
        for _ in range(12):
            pipe.step_resident(pd_, gd_, cd, assume_ready=True)
        torch.cuda.synchronize()

        def timed(fn, n=20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            k0 = L.kernel_launches()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            th = time.perf_counter() - t0
            torch.cuda.synchronize()
            return {"host_ms": round(1e3 * th / n, 3), "device_ms": round(e0.elapsed_time(e1) / n, 3),
                    "eager_library_kernels": (L.kernel_launches() - k0) / n}

        nmalloc = lambda: int(torch.cuda.memory_stats(dev).get("num_device_alloc", 0))
        m0 = nmalloc()
        out = {"step": timed(lambda: pipe.step_resident(pd_, gd_, cd, assume_ready=True))}
        out["step"]["cudaMalloc_calls"] = nmalloc() - m0
        out["step"]["gc_so_far"] = dict(GC_STATS, t0=None)
        trace = []
        for _ in range(30):
            t0 = time.perf_counter()
            pipe.step_resident(pd_, gd_, cd, assume_ready=True)
            trace.append(round(1e3 * (time.perf_counter() - t0), 2))
        torch.cuda.synchronize()
        out["host_ms_per_step_trace"] = trace
        import cProfile, pstats, io
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(20):
            pipe.step_resident(pd_, gd_, cd, assume_ready=True)
        torch.cuda.synchronize()
        pr.disable()
        sio = io.StringIO()
        pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(8)
        print(sio.getvalue()[:2500])
        out["transform_alone"] = timed(lambda: pipe.transform(pd_, gd_))
        pat, gr = pipe.transform(pd_, gd_)
        for x in (pat, gr):
            x.csr_in, x.csr_out
        out["train_alone"] = timed(lambda: pipe.train_on(pat, gr, cd))
        out["replayed_library_kernels_per_step"] = pipe.replayed_library_kernels() / max(1, getattr(pipe, "_replays", 1))
        print(key, out, flush=True)
        pipe._graphs.clear()


if __name__ == "__main__":
    main()
