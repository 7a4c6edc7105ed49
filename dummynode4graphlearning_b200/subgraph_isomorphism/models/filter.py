"""Label filter (``subgraph_isomorphism/models/filter.py:6-16``).

``ScalarFilter.forward(p_x, g_x)`` keeps the reference's dense signature for callers that hold padded
label matrices; the models themselves call ``gate_from_graphs`` which evaluates the same predicate with
one CUDA kernel on the ragged batch (no (B, Lg, Lp) temporary, no per-graph Python slicing --
basemodel.py:830-847)."""
import torch as th
import torch.nn as nn

from ... import ops


class ScalarFilter(nn.Module):
    def forward(self, p_x, g_x):
        """p_x: bsz x l1 (x1), g_x: bsz x l2 (x1) -> bsz x l2 (x1) bool"""
        return ((g_x.unsqueeze(2) - p_x.unsqueeze(1)) == 0).any(dim=2)

    def gate_from_graphs(self, pattern, graph, kind="node"):
        Lp_max = pattern.padded_num_nodes() if kind == "node" else pattern.padded_num_edges()
        return ops.label_filter_gate(graph, pattern, Lp_max, kind)
