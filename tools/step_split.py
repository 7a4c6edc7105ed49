"""How long do the two halves of the C2 step take on their own?  (a) the CUDA-graph replay of the train step alone,
(b) the transform alone (host time and time to completion), (c) both overlapped as in ClassificationPipeline."""
import os, sys, time
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
torch.manual_seed(0)
model = GIN(args).to(dev)
pipe = ClassificationPipeline(model, FlatAdam(model.parameters(), lr=0.01), mode="conj", num_node_labels=2, node_label_min=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(30):
    pipe.step_resident(dev_batch, assume_ready=True)
torch.cuda.synchronize()
K = 100
data = pipe.transform(dev_batch)
torch.cuda.synchronize()
for with_flush in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        if with_flush:
            flush.fill_(1)
        pipe.train_on(data)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print("replay only%s: device %.3f ms/step, host %.3f ms/step" % (" + flush" if with_flush else "", e0.elapsed_time(e1) / K, 1e3 * (t1 - t0) / K))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K):
    d = pipe.transform(dev_batch)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("transform only (one stream): device %.3f ms/step, host %.3f ms/step" % (e0.elapsed_time(e1) / K, 1e3 * (t1 - t0) / K))
for with_flush in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        if with_flush:
            flush.fill_(1)
        pipe.step_resident(dev_batch, assume_ready=True)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("full step%s: device %.3f ms/step, host %.3f ms/step (with final sync %.3f)" % (" + flush" if with_flush else "", e0.elapsed_time(e1) / K, 1e3 * (t1 - t0) / K, 1e3 * (t2 - t0) / K))
pipe.overlap = False
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
for _ in range(K):
    pipe.step_resident(dev_batch, assume_ready=True)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("full step, one stream: device %.3f ms/step, host %.3f ms/step" % (e0.elapsed_time(e1) / K, 1e3 * (t1 - t0) / K))
