"""cProfile of the HOST side of the C2 transform (the part of the step that is not a CUDA graph): which Python-level calls
the ~1.1 ms go to.  python tools/host_profile_transform.py"""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from dummynode4graphlearning_b200 import synth, transforms as T
from dummynode4graphlearning_b200.graph_classification.models import GIN
from dummynode4graphlearning_b200.optim import FlatAdam
from dummynode4graphlearning_b200.pipelines import ClassificationPipeline

dev = torch.device("cuda:0")
raw = synth.tu_batch("proteins", 1113, seed=0)
dev_batch = T.to_device({k: v for k, v in raw.items() if k != "vattr"}, dev)
args = Namespace(num_features=2, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                 additional={"train_eps": True, "num_layers": 4, "aggregation": "sum"}, epochs=1, device=str(dev))
model = GIN(args).to(dev)
pipe = ClassificationPipeline(model, FlatAdam(model.parameters(), lr=0.01), mode="conj", num_node_labels=2, node_label_min=0)
for _ in range(5):
    pipe.transform(dev_batch)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); d = pipe.transform(dev_batch); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
print("transform: host %.3f ms, +sync %.3f ms (median of 20)" % (1e3 * sorted(t[0] for t in ts)[10], 1e3 * sorted(t[1] for t in ts)[10]))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    d = pipe.transform(dev_batch)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)

# ---- the whole step (transform on the second stream + CUDA-graph replay): where does the host time go?
for _ in range(12):
    pipe.step_resident(dev_batch, assume_ready=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    pipe.step_resident(dev_batch, assume_ready=True)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("step_resident x50: host %.3f ms/step, with final sync %.3f ms/step" % (1e3 * (t1 - t0) / 50, 1e3 * (t2 - t0) / 50))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    pipe.step_resident(dev_batch, assume_ready=True)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
st.sort_stats("cumulative").print_stats(22)
