"""per-step timeline of ClassificationPipeline.step_resident / step on the C2 batch (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
import torch
from dummynode4graphlearning_b200 import synth, transforms as T
from dummynode4graphlearning_b200.graph_classification.models import GIN
from dummynode4graphlearning_b200.pipelines import ClassificationPipeline, pin_batch, _CapturedStep

dev = torch.device("cuda:0")
raw = synth.tu_batch("proteins", 1113, seed=0)
host = pin_batch({k: v for k, v in raw.items() if k != "vattr"})
dev_batch = T.to_device({k: v for k, v in raw.items() if k != "vattr"}, dev)
torch.manual_seed(0)
args = Namespace(num_features=4, hidden_dim=32, num_classes=2, dropout_ratio=0.0,
                 additional={"train_eps": True, "num_layers": 4, "aggregation": "sum"}, epochs=1, device=str(dev))
model = GIN(args).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.01, capturable=True)
pipe = ClassificationPipeline(model, opt, mode="conj", num_node_labels=4, node_label_min=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def report(tag):
    ents = [e for e in pipe._graphs.values()]
    print(tag, "graphs:", len(ents), "replays:", [e.replays if isinstance(e, _CapturedStep) else e for e in ents], flush=True)


for phase in ("resident", "host", "resident", "host"):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(26)]
    wall = []
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(25):
        t0 = time.perf_counter()
        flush.fill_(1)
        if phase == "resident":
            pipe.step_resident(dev_batch)
        else:
            pipe.step(host)
        wall.append((time.perf_counter() - t0) * 1e3)
        ev[i + 1].record()
    torch.cuda.synchronize()
    gpu = [ev[i].elapsed_time(ev[i + 1]) for i in range(25)]
    print(phase, "gpu ms:", " ".join("%.2f" % x for x in gpu))
    print(phase, "host ms:", " ".join("%.2f" % x for x in wall))
    report(phase)
