"""per-step loss of the C4 / C3 counting step under torch.optim.AdamW vs FlatAdam (eager, serial)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from bench_counting import CFG, build
from dummynode4graphlearning_b200 import synth, transforms as T
from dummynode4graphlearning_b200.optim import FlatAdam
from dummynode4graphlearning_b200.pipelines import CountingPipeline
dev = torch.device("cuda:0")
for cfgname in ("c4", "c3"):
    name, shape, bs, over = CFG[cfgname]
    p, g, counts = synth.counting_batch(shape, bs, seed=0)
    pd_, gd_, cd = T.to_device(p, dev), T.to_device(g, dev), torch.from_numpy(counts).to(dev)
    res = {}
    for opt_name in ("torch", "flat", "flat_graph"):
        model, cfg, kw = build(name, shape, over, dev)
        if opt_name == "torch":
            opt = torch.optim.AdamW(model.parameters(), lr=1e-3, amsgrad=True)
        else:
            opt = FlatAdam(model.parameters(), lr=1e-3, weight_decay=1e-2, amsgrad=True, decoupled_weight_decay=True)
        pipe = CountingPipeline(model, opt, cfg, add_dummy=True, rep_reg_w=1e-3, cuda_graphs=(opt_name == "flat_graph"), overlap=False)
        res[opt_name] = [float(pipe.step_resident(pd_, gd_, cd).item()) for _ in range(40)]
    for k, v in res.items():
        print(cfgname, k, " ".join("%.4g" % x for x in v[:12]), "...", " ".join("%.4g" % x for x in v[-3:]))
