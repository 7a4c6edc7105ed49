"""End-to-end public API: host graph batches in, loss out.

``ClassificationPipeline.step(host_batch)`` is the call a user of the classification side makes per
mini-batch: pinned host arrays of RAW TU-shaped graphs -> H2D -> dummy augmentation -> edge-to-vertex
transform -> PyG canonicalisation -> CSR build -> GIN forward + nll_loss + backward -> (gradient all-reduce)
-> Adam step -> loss read back.  Everything between the two copies runs on the GPU; it replaces the
reference's offline ``tu_data_processing.py`` + ``PYGDataset`` + ``main.py:train`` loop body
(tu_data_processing.py:417-455, graph_neural_networks/dataset.py:141-168, main.py:37-43).

``CountingPipeline.step(pattern, graph, counts)`` does the same for the subgraph-counting side
(train.py:1322-1334 augmentation + train_epoch body :753-838).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import transforms as T
from .graph import BatchedGraph
from .graph_classification.data import Batch
from .parallel import GradientBucket, is_distributed

_HOST_KEYS = ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "vid", "eid", "y", "vattr", "e_is_reversed")


def pin_batch(b):
    """numpy batch dict -> pinned host tensors (int32 / float32 / int64 for y)."""
    out = {}
    for k, v in b.items():
        if k in _HOST_KEYS:
            t = torch.as_tensor(np.ascontiguousarray(v))
            out[k] = t.pin_memory() if torch.cuda.is_available() else t
        else:
            out[k] = v
    return out


def host_bytes(b):
    return int(sum(v.numel() * v.element_size() for v in b.values() if isinstance(v, torch.Tensor)))


def upload(b, device):
    out = {}
    for k, v in b.items():
        out[k] = v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else v
    return out


class ClassificationPipeline:
    def __init__(self, model, optimizer, mode="conj", num_node_labels=None, num_edge_labels=None,
                 node_label_min=None, with_edge_attr=False):
        """mode: 'dummy' (DUMMY_ graphs), 'conj' (CONJ_: dummy + edge-to-vertex), 'line' (LINE_), 'raw'."""
        self.model, self.opt, self.mode = model, optimizer, mode
        self.nvl, self.nel = num_node_labels, num_edge_labels
        self.node_label_min, self.with_edge_attr = node_label_min, with_edge_attr
        self.bucket = GradientBucket(model.parameters())
        self.device = next(model.parameters()).device
        self.global_batch = None

    def transform(self, dev_batch):
        """raw TU-shaped device batch -> PyG-style Batch with compiled structure."""
        b = dev_batch
        if self.mode in ("dummy", "conj"):
            b = T.tu_add_dummy(b)
        if self.mode in ("conj", "line"):
            b = T.tu_conjugate(b)
            b.pop("eattr", None)
        b["has_edge_labels"] = True
        # GIN never reads edge_attr (gconv.py:204); RGIN does (rgconv.py:109-111)
        can = T.pyg_canonicalize(b, self.nvl, self.nel, node_label_min=self.node_label_min,
                                 with_edge_attr=self.with_edge_attr)
        data = Batch.from_canonical(can)
        data.structure  # compile the CSR pair now (part of the transform cost)
        return data

    def train_on(self, data):
        self.model.train()
        self.bucket.zero()
        out = self.model(data)
        loss = F.nll_loss(out, data.y)                      # main.py:41
        loss.backward()
        if is_distributed():
            gb = self.global_batch or data.num_graphs * torch.distributed.get_world_size()
            self.bucket.all_reduce(data.num_graphs / gb)
        self.opt.step()                                     # main.py:43
        return loss

    def step_resident(self, dev_batch):
        """inputs already in HBM: transform + train step; returns the loss tensor (no host sync)."""
        return self.train_on(self.transform(dev_batch))

    def step(self, host_batch):
        """host buffers in, python float out: H2D + transform + train step + D2H of the loss."""
        dev = upload(host_batch, self.device)
        dev = {k: (v.to(torch.int32) if isinstance(v, torch.Tensor) and v.dtype == torch.int64 and k != "y" else v)
               for k, v in dev.items()}
        return float(self.step_resident(dev).item())


class CountingPipeline:
    def __init__(self, model, optimizer, config, add_dummy=True, rep_reg_w=0.0, neg_slp=0.01, max_grad_norm=8.0):
        """config: the dataset maxima BEFORE augmentation (max_npv ... max_ngel)."""
        self.model, self.opt, self.cfg, self.add_dummy = model, optimizer, config, add_dummy
        self.rep_reg_w, self.neg_slp, self.max_grad_norm = rep_reg_w, neg_slp, max_grad_norm
        self.bucket = GradientBucket(model.parameters())
        self.device = next(model.parameters()).device
        self.global_batch = None

    def transform(self, p_dev, g_dev):
        c = self.cfg
        if self.add_dummy:
            p_dev = T.sub_add_dummy(p_dev, c["max_npv"], c["max_npvl"], c["max_npe"], c["max_npel"])
            g_dev = T.sub_add_dummy(g_dev, c["max_ngv"], c["max_ngvl"], c["max_nge"], c["max_ngel"])
        pattern, graph = BatchedGraph.from_batch(p_dev, self.device), BatchedGraph.from_batch(g_dev, self.device)
        for g in (pattern, graph):   # compile both CSRs + degrees now (calculate_degrees, train.py:1355-1356)
            g.in_degrees()
            g.out_degrees()
        return pattern, graph

    def loss_fn(self, out, counts):
        crit = lambda pred, target, slp: F.mse_loss(F.leaky_relu(pred, slp), target)   # train.py:624-625
        loss = crit(out["pred_c"], counts.float().view(-1, 1), self.neg_slp)
        if self.rep_reg_w:
            for k in ("p_v_rep", "p_e_rep", "g_v_rep", "g_e_rep"):                      # train.py:801-811
                if out[k] is not None:
                    loss = loss + self.rep_reg_w * crit(out[k], torch.zeros_like(out[k]), 1) * out[k].size(1)
        return loss

    def train_on(self, pattern, graph, counts):
        self.model.train()
        self.bucket.zero()
        out = self.model(pattern, graph)
        loss = self.loss_fn(out, counts)
        loss.backward()
        if is_distributed():
            gb = self.global_batch or pattern.batch_size * torch.distributed.get_world_size()
            self.bucket.all_reduce(pattern.batch_size / gb)
        if self.max_grad_norm and self.max_grad_norm > 0:   # clip AFTER the reduction (train.py:833-834)
            torch.nn.utils.clip_grad_norm_(self.bucket.params, self.max_grad_norm)
        self.opt.step()
        return loss

    def step_resident(self, p_dev, g_dev, counts_dev):
        pattern, graph = self.transform(p_dev, g_dev)
        return self.train_on(pattern, graph, counts_dev)

    def step(self, p_host, g_host, counts_host):
        p_dev, g_dev = upload(p_host, self.device), upload(g_host, self.device)
        counts = counts_host.to(self.device, non_blocking=True)
        return float(self.step_resident(p_dev, g_dev, counts).item())
