"""where does the HOST time of one C2 step go?  torch.profiler over 3 steps (CPU + CUDA activities)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from torch.profiler import ProfilerActivity, profile
from dummynode4graphlearning_b200 import synth, transforms as T
from dummynode4graphlearning_b200.graph_classification.models import GIN
from dummynode4graphlearning_b200.pipelines import ClassificationPipeline

dev = torch.device("cuda:0")
raw = synth.tu_batch("proteins", 1113, seed=0)
dev_batch = T.to_device({k: v for k, v in raw.items() if k != "vattr"}, dev)
torch.manual_seed(0)
args = Namespace(num_features=4, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                 additional={"train_eps": True, "num_layers": 4, "aggregation": "sum"}, epochs=1, device=str(dev))
model = GIN(args).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.01)
pipe = ClassificationPipeline(model, opt, mode="conj", num_node_labels=4, node_label_min=0)
for _ in range(5):
    pipe.step_resident(dev_batch)
torch.cuda.synchronize()
# wall-clock split: transform vs train, host time only (no sync inside) and with sync
for name, fn in (("transform", lambda: pipe.transform(dev_batch)),):
    t0 = time.perf_counter(); d = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%s: host %.3f ms, +sync %.3f ms" % (name, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))
data = pipe.transform(dev_batch)
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); pipe.train_on(data); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("train_on: host %.3f ms, +sync %.3f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
model.train()
for i in range(2):
    t0 = time.perf_counter(); out = model(data); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    loss = torch.nn.functional.nll_loss(out, data.y)
    t3 = time.perf_counter(); loss.backward(); t4 = time.perf_counter(); torch.cuda.synchronize(); t5 = time.perf_counter()
    t6 = time.perf_counter(); opt.step(); t7 = time.perf_counter(); torch.cuda.synchronize()
    print("fwd host %.3f (+sync %.3f)  bwd host %.3f (+sync %.3f)  opt host %.3f" % (1e3*(t1-t0), 1e3*(t2-t0), 1e3*(t4-t3), 1e3*(t5-t3), 1e3*(t7-t6)))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        pipe.step_resident(dev_batch)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))
