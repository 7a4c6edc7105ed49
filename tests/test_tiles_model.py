"""CPU: invariants of the aggregation kernel's tile construction on the host model (tools/k1_tiles_model.py restates
csrc/spmm_tiled.cu tile_boundary / make_row_tiles_kernel / collect_cut_heavy_kernel): for the shipped rule and for the
prepared whole-graph rule (-DDN4GL_TILE_WHOLE_GRAPHS) the tiles partition [0, N), closed tiles hold whole graphs and fit
one stage, and the +-1 window search finds every row's tile.  The device code itself is covered by the GPU aggregation
tests (tests/test_agg_gpu.py, every ring size)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("k1_tiles_model", os.path.join(ROOT, "tools", "k1_tiles_model.py"))
model = importlib.util.module_from_spec(spec)
spec.loader.exec_module(model)


@pytest.mark.parametrize("C", [1, 2, 7, 64, 311])
def test_tiles_partition_rows_under_both_rules(C):
    rng = np.random.default_rng(C)
    for trial in range(120):
        n = rng.integers(1, 4 * C + 2, int(rng.integers(1, 40)))
        if trial % 3 == 0:
            n = np.minimum(n, int(rng.integers(1, 2 * C + 2)))
        seg = np.concatenate([[0], np.cumsum(n)])
        cut_old, _ = model.check(seg, C, whole=False)
        cut_new, biggest = model.check(seg, C, whole=True)
        assert cut_new <= cut_old + 1e-12        # the whole-graph rule never cuts more rows than the shipped one
        assert biggest <= 2 * C


def test_whole_graph_rule_keeps_a_spanning_graph_in_one_tile():
    # window 10; graphs of 8, 19 (spans window 1 = rows [10, 20)) and 3 rows: the shipped rule cuts the 19-row graph at 10
    seg = np.array([0, 8, 27, 30])
    assert model.make_tiles(seg, 10, whole=False) == [(0, 10, True), (10, 27, True), (27, 30, False)]
    assert model.make_tiles(seg, 10, whole=True) == [(0, 8, False), (8, 27, False), (27, 30, False)]
    # a graph longer than two windows is still cut
    seg = np.array([0, 4, 29, 33])
    assert any(cut for _, _, cut in model.make_tiles(seg, 10, whole=True))
