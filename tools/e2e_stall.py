"""Which part of ClassificationPipeline.step_async stalls when a step takes 5-100 ms instead of 1.2 ms?  Runs the bench's
software-pipelined e2e loop for many steps with a host timer around every phase and prints the outliers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from dummynode4graphlearning_b200 import pipelines as P, synth
from dummynode4graphlearning_b200.graph_classification.models import GIN
from dummynode4graphlearning_b200.optim import FlatAdam

dev = torch.device("cuda:0")
raw = synth.tu_batch("proteins", 1113, seed=0)
host = P.pin_batch({k: v for k, v in raw.items() if k != "vattr"})
args = Namespace(num_features=2, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                 additional={"train_eps": True, "num_layers": 4, "aggregation": "sum"}, epochs=1, device=str(dev))
torch.manual_seed(0)
model = GIN(args).to(dev)
pipe = P.ClassificationPipeline(model, FlatAdam(model.parameters(), lr=0.01), mode="conj", num_node_labels=2, node_label_min=0)
phases = {}
cur = {}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            cur[name] = cur.get(name, 0.0) + time.perf_counter() - t0
    return w


pipe._upload = timed("upload", pipe._upload)
pipe.transform = timed("transform", pipe.transform)
pipe._hand_over = timed("hand_over", pipe._hand_over)
pipe.train_on = timed("train_on", pipe.train_on)
pipe._throttle = timed("throttle", pipe._throttle)
import gc
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
pending = None
for _ in range(40):
    nxt = pipe.step_async(host)
    if pending is not None:
        pending.result()
    pending = nxt
pending.result()
gc.collect(); gc.freeze()
torch.cuda.synchronize()
N = int(os.environ.get("STEPS", "3000"))
rows = []
pending = None
t_prev = time.perf_counter()
for i in range(N):
    cur.clear()
    flush.fill_(1)
    nxt = pipe.step_async(host)
    t1 = time.perf_counter()
    if pending is not None:
        pending.result()
    pending = nxt
    t2 = time.perf_counter()
    rows.append((t2 - t_prev, t2 - t1, dict(cur)))
    t_prev = t2
tot = sorted(r[0] for r in rows)
print("steps %d  median %.3f ms  mean %.3f ms  p99 %.3f ms  max %.3f ms" % (N, 1e3 * tot[N // 2], 1e3 * sum(tot) / N, 1e3 * tot[int(N * 0.99)], 1e3 * tot[-1]))
print("cudaMalloc calls:", torch.cuda.memory_stats(dev).get("num_device_alloc"))
for i, (t, tr, ph) in enumerate(rows):
    if t > 3e-3:
        print("step %d: %.2f ms  result %.2f  " % (i, 1e3 * t, 1e3 * tr) + " ".join("%s %.2f" % (k, 1e3 * v) for k, v in ph.items()))
