"""CPU: the on-disk TU reader / writer against the reference-generated goldens, and round trips."""
import os
import tempfile

import numpy as np
import pytest

from dummynode4graphlearning_b200.graph_classification import io as tuio
from helpers import load_golden
from oracle import transforms as OT


@pytest.mark.parametrize("case", ["appB", "mutag12", "proteins8"])
def test_writer_matches_reference_save_graph_data(case):
    """tu_file_lines(CONJ batch) == the text the reference's save_graph_data wrote for the same graphs."""
    g = load_golden("transforms.pt")["tu/" + case]
    conj = OT.tu_conjugate(OT.tu_add_dummy(g["inp"]))
    mine = tuio.tu_file_lines(conj)
    for suffix, lines in g["conj_files"].items():
        assert mine[suffix] == lines, suffix


@pytest.mark.parametrize("case", ["mutag12", "proteins8"])
def test_reader_round_trip_and_reference_loader_equivalence(case):
    """write (reference layout) -> load_tu_dir -> the same batch; and dummy-augmenting the loaded batch gives what the
    reference's load_graph_data_from_TUDatadir(with_dummy=True) built from those files (golden)."""
    g = load_golden("transforms.pt")["tu/" + case]
    b = dict(g["inp"])
    b["vid"] = np.concatenate([np.arange(n) for n in np.diff(b["node_ptr"])]).astype(np.int32)
    b["eid"] = np.concatenate([np.arange(n) for n in np.diff(b["edge_ptr"])]).astype(np.int32)
    with tempfile.TemporaryDirectory() as d:
        raw = os.path.join(d, "DS", "raw")
        tuio.save_tu_dir(b, raw)
        assert sorted(os.listdir(raw))[0].startswith("DS_")
        back = tuio.load_tu_dir(os.path.join(d, "DS"))
    for k in ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel"):
        assert np.array_equal(back[k], b[k]), k
    if "vattr" in b:
        assert np.allclose(back["vattr"], b["vattr"], rtol=0, atol=0)
    dummy = OT.tu_add_dummy(back)
    for k in ("node_ptr", "edge_ptr", "src", "dst", "vlabel", "elabel", "v_is_dummy", "e_is_dummy"):
        assert np.array_equal(dummy[k], g["dummy"][k]), k


def test_label_shift_and_defaults():
    """labels whose minimum is not 1 are shifted to start at 1; missing label files mean all ones (:153-171)."""
    with tempfile.TemporaryDirectory() as d:
        def w(name, lines):
            with open(os.path.join(d, "X_" + name + ".txt"), "w") as f:
                f.write("\n".join(lines) + "\n")
        w("A", ["1, 2", "2, 1", "3, 4", "4, 3"])
        w("graph_indicator", ["1", "1", "2", "2"])
        w("node_labels", ["0", "3", "0", "1"])
        w("graph_labels", ["-1", "1"])
        b = tuio.load_tu_dir(d)
    assert b["vlabel"].tolist() == [1, 4, 1, 2] and b["elabel"].tolist() == [1, 1, 1, 1] and not b["has_edge_labels"]
    assert b["node_ptr"].tolist() == [0, 2, 4] and b["edge_ptr"].tolist() == [0, 2, 4] and b["y"].tolist() == [-1, 1]
    assert b["src"].tolist() == [0, 1, 2, 3] and b["dst"].tolist() == [1, 0, 3, 2]
