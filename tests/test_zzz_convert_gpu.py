"""GPU: graph_classification.io.convert_tu_dataset -- the offline DUMMY_ / LINE_ / CONJ_ conversion of a TU directory
(tu_data_processing.py __main__, :431-456) with the graph construction on the GPU -- against the oracle transforms +
the writer, file by file.  (The host logic and the writer are pinned to the reference in tests/test_tu_io.py.)"""
import os
import tempfile

import numpy as np
import pytest

from dummynode4graphlearning_b200.graph_classification import io as tuio
from oracle import transforms as OT
from test_tu_io import _tu_dir_with_decimal_attributes

pytestmark = pytest.mark.gpu


def test_convert_tu_dataset_on_gpu_matches_oracle(device):
    with tempfile.TemporaryDirectory() as d:
        raw = _tu_dir_with_decimal_attributes(d, nb=40, seed=8)
        out = tuio.convert_tu_dataset(raw, "PROTEINS", str(device))
        src = tuio.load_tu_dir(raw)
        dummy = OT.tu_add_dummy(src)
        expect = {"DUMMY_": dummy, "LINE_": OT.tu_conjugate(src), "CONJ_": OT.tu_conjugate(dummy)}
        for prefix, ref in expect.items():
            got = {n.split("PROTEINS_", 1)[1][:-4]: open(os.path.join(out[prefix], n)).read().split("\n")[:-1]
                   for n in os.listdir(out[prefix])}
            lines = tuio.tu_file_lines(dict(ref, y=src["y"]))
            for suffix in ("graph_indicator", "A", "node_labels", "edge_labels", "node_ids", "edge_ids", "graph_labels"):
                assert got[suffix] == lines[suffix], (prefix, suffix)
            # attributes travel on the host in float64: compare as numbers with the oracle's (float32) passthrough
            for suffix in ("node_attributes", "edge_attributes"):
                if suffix in lines:
                    assert suffix in got, (prefix, suffix)
                    assert np.allclose(np.array(got[suffix], dtype=np.float64), np.array(lines[suffix], dtype=np.float64),
                                       rtol=1e-6, atol=1e-7), (prefix, suffix)
