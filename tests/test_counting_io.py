"""CPU: the GML + CSV reader of the counting data sets (subgraph_isomorphism/utils/io.py:43-220):
known-answer GML texts, round trips through the reference's directory layout, edge cases, and -- where /root/reference
exists -- equality with the reference's own ``load_data`` run under the shims."""
import os
import tempfile

import numpy as np
import pytest

from dummynode4graphlearning_b200 import synth
from dummynode4graphlearning_b200.subgraph_isomorphism import io as sio

# what python-igraph 0.9 writes for a 3-node directed multigraph (brackets on their own lines, numeric attributes bare)
IGRAPH_GML = """Creator "igraph version 0.9.11 Mon Oct  3 2022"
Version 1
graph
[
  directed 1
  node
  [
    id 0
    label 2
  ]
  node
  [
    id 1
    label 0
  ]
  node
  [
    id 2
    label 1
  ]
  edge
  [
    source 0
    target 1
    label 3
    key 0
  ]
  edge
  [
    source 0
    target 1
    label 1
    key 1
  ]
  edge
  [
    source 2
    target 0
    label 0
    key 0
  ]
]
"""

# the same graph as another GML producer would lay it out: one-line lists, comment, float-typed numbers, ids that are
# neither 0-based nor in order, a string attribute and a nested list igraph ignores
COMPACT_GML = """# produced elsewhere
graph [ directed 1 name "toy"
  node [ id 10 label 2.0 graphics [ x 1.5 y 2 ] ]
  node [ id 7 label 0 ] node [ id 42 label 1 tag "a b ] c" ]
  edge [ source 10 target 7 label 3 key 0 ] edge [ source 10 target 7 label 1 key 1 ]
  edge [ source 42 target 10 label 0.0 key 0 ]
]
"""


def _write(d, name, text):
    p = os.path.join(d, name)
    with open(p, "w") as f:
        f.write(text)
    return p


def test_known_answer_igraph_layout():
    with tempfile.TemporaryDirectory() as d:
        g = sio.read_gml_graph(_write(d, "g.gml", IGRAPH_GML))
    assert g["num_nodes"] == 3
    assert g["src"].tolist() == [0, 0, 2] and g["dst"].tolist() == [1, 1, 0]
    assert g["vid"].tolist() == [0, 1, 2] and g["vlabel"].tolist() == [2, 0, 1]
    assert g["elabel"].tolist() == [3, 1, 0] and g["ekey"].tolist() == [0, 1, 0]
    assert all(g[k].dtype == np.int64 for k in ("src", "dst", "vid", "vlabel", "elabel", "ekey"))


def test_known_answer_compact_layout_ids_resolve_through_node_id():
    with tempfile.TemporaryDirectory() as d:
        g = sio.read_gml_graph(_write(d, "g.gml", COMPACT_GML))
    assert g["src"].tolist() == [0, 0, 2] and g["dst"].tolist() == [1, 1, 0]      # positions, not the ids 10 / 7 / 42
    assert g["vid"].tolist() == [10, 7, 42] and g["vlabel"].tolist() == [2, 0, 1]
    assert g["elabel"].tolist() == [3, 1, 0] and g["ekey"].tolist() == [0, 1, 0]
    p = sio.parse_gml(COMPACT_GML)
    assert p["directed"] is True and p["vattr"]["tag"] == [None, None, "a b ] c"] and "graphics" not in p["vattr"]


def test_edge_cases():
    with tempfile.TemporaryDirectory() as d:
        g = sio.read_gml_graph(_write(d, "e.gml", "graph [ directed 1 node [ id 0 label 5 ] ]"))
        assert g["num_nodes"] == 1 and g["src"].shape == (0,) and g["elabel"].shape == (0,) and g["vlabel"].tolist() == [5]
        g = sio.read_gml_graph(_write(d, "z.gml", "graph [ directed 1 ]"))
        assert g["num_nodes"] == 0 and g["src"].shape == (0,) and g["vid"].shape == (0,)
        for bad in ("graph [ node [ id 0 label 1 ] node [ id 0 label 2 ] ]",                      # duplicate id
                    "graph [ node [ id 0 label 1 ] edge [ source 0 target 9 label 0 key 0 ] ]",   # unknown endpoint
                    "graph [ node [ id 0 label 1 ]",                                               # unterminated
                    "graph [ node [ id 0 ] ]",                                                     # no label
                    "node [ id 0 label 1 ]"):                                                      # no graph block
            with pytest.raises(ValueError):
                sio.read_gml_graph(_write(d, "b.gml", bad))


def _dump(d, layout, seed=7, B=23):
    from oracle import ref_drive as rd        # writer helper only (test scaffolding)
    p, g, c = synth.counting_batch("small", B, seed=seed)
    mats = synth.random_subisomorphisms(p, g, seed=seed)
    return (p, g, c, mats) + tuple(rd.write_counting_dirs(d, p, g, c, mats, layout))


def test_round_trip_own_graph_layout():
    """batch -> patterns/ graphs/P_i/ metadata/ -> load_data -> collate gives the batch back (all three splits)."""
    with tempfile.TemporaryDirectory() as d:
        p, g, c, mats, pd, gd, md = _dump(d, "own")
        data, shared = sio.load_data(pd, gd, md, num_workers=1)
        data_mt, _ = sio.load_data(pd, gd, md, num_workers=4)
    assert not shared
    assert [[x["id"] for x in data[k]] for k in data] == [[x["id"] for x in data_mt[k]] for k in data_mt]
    samples = sorted((x for k in data for x in data[k]), key=lambda x: int(x["id"].rsplit("_", 1)[-1]))
    assert len(samples) == 23
    for x in samples:                                  # G_i_i: i % 10 decides the split
        i = int(x["id"].rsplit("_", 1)[-1])
        split = "dev" if i % 10 == 0 else "test" if i % 10 == 1 else "train"
        assert any(x is y for y in data[split])
    pb, gb, counts, ms = sio.collate(samples)
    for k in ("node_ptr", "edge_ptr", "src", "dst", "vid", "vlabel", "eid", "elabel"):
        assert np.array_equal(pb[k], p[k]) and pb[k].dtype == p[k].dtype, k
        assert np.array_equal(gb[k], g[k]) and gb[k].dtype == g[k].dtype, k
    assert np.array_equal(counts, c)
    assert all(np.array_equal(a, b) and a.shape == b.shape for a, b in zip(ms, mats))


def test_shared_graph_layout_and_index_files():
    with tempfile.TemporaryDirectory() as d:
        p, g, c, mats, pd, gd, md = _dump(d, "shared", B=10)
        data, shared = sio.load_data(pd, gd, md, num_workers=1)
        assert shared
        ids = {k: sorted(x["id"] for x in data[k]) for k in data}
        assert ids["train"] == sorted("P_%d-G_%d" % (a, b) for a in range(3) for b in range(10) if b % 3 == 2)
        assert ids["dev"] == sorted("P_%d-G_%d" % (a, b) for a in range(3) for b in range(10) if b % 3 == 0)
        assert ids["test"] == sorted("P_%d-G_%d" % (a, b) for a in range(3) for b in range(10) if b % 3 == 1)
        _write(md, "train.txt", "0\n1\n2\n")              # explicit index files override the modulo rule, per split
        _write(md, "test.txt", "9\n")
        data, _ = sio.load_data(pd, gd, md, num_workers=1)
        assert sorted(x["id"] for x in data["train"]) == sorted("P_%d-G_%d" % (a, b) for a in range(3) for b in range(3))
        assert sorted(x["id"] for x in data["test"]) == ["P_0-G_9", "P_1-G_9", "P_2-G_9"]
        assert sorted(x["id"] for x in data["dev"]) == ids["dev"]
    x = next(x for x in data["train"] if x["id"] == "P_1-G_1")
    assert x["counts"] == int(c[1]) and np.array_equal(x["subisomorphisms"].reshape(-1, mats[1].shape[1]), mats[1])


def test_metadata_cell_formats():
    with tempfile.TemporaryDirectory() as d:
        _write(d, "P_0.csv", 'g_id,counts,subisomorphisms\nG_0,2,"[[0, 1], [2, 3]]"\nG_1,0,[]\nG_2,1,"[(4, 5)]"\n')
        m = sio.read_metadata_from_dir(d, num_workers=1)
    assert m["P_0"]["G_0"]["subisomorphisms"].tolist() == [[0, 1], [2, 3]] and m["P_0"]["G_0"]["counts"] == 2
    assert m["P_0"]["G_1"]["subisomorphisms"].shape == (0,) and m["P_0"]["G_1"]["subisomorphisms"].dtype == np.int64
    assert m["P_0"]["G_2"]["subisomorphisms"].tolist() == [[4, 5]]


@pytest.mark.reference_live
@pytest.mark.parametrize("layout,seed", [("own", 11), ("shared", 12)])
def test_load_data_equals_reference_live(layout, seed):
    """the unmodified utils/io.py:load_data (igraph.read = oracle/shims' line-oriented GML reader, written
    independently of the product's tokenizer) returns the same samples, in the same order, as this loader."""
    from oracle import ref_drive as rd
    with tempfile.TemporaryDirectory() as d:
        _, _, _, _, pd, gd, md = _dump(d, layout, seed=seed, B=14)
        if layout == "shared":
            _write(md, "dev.txt", "3\n4\n")
        mine, sh = sio.load_data(pd, gd, md, num_workers=1)
        ref, rsh = rd.ref_load_data(pd, gd, md)
    assert sh == rsh
    for split in ("train", "dev", "test"):
        assert [x["id"] for x in mine[split]] == [x["id"] for x in ref[split]], split
        for a, b in zip(mine[split], ref[split]):
            assert a["counts"] == b["counts"]
            assert np.array_equal(a["subisomorphisms"], b["subisomorphisms"])
            for side in ("pattern", "graph"):
                assert a[side]["num_nodes"] == b[side]["num_nodes"]
                for k in ("src", "dst", "vid", "vlabel", "elabel", "ekey"):
                    assert np.array_equal(a[side][k], b[side][k]), (split, a["id"], side, k)


def _make_tree(root):
    for rel in ("a/x/1.gml", "a/x/2.gml", "a/y.gml", "b/3.gml", "c/d/e/4.gml", "top.gml"):
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        open(p, "w").close()
    os.makedirs(os.path.join(root, "empty"))


def test_directory_walks_known_answer():
    """get_subdirs: children before their parent, leaves only by default; get_files: every file exactly once."""
    with tempfile.TemporaryDirectory() as d:
        _make_tree(d)
        rel = lambda ps: [os.path.relpath(p, d) for p in ps]
        leaves, every, files = rel(sio.get_subdirs(d)), rel(sio.get_subdirs(d, leaf_only=False)), rel(sio.get_files(d))
    assert sorted(leaves) == ["a/x", "b", "c/d/e", "empty"]
    assert sorted(every) == [".", "a", "a/x", "b", "c", "c/d", "c/d/e", "empty"] and every[-1] == "."
    for child, parent in (("a/x", "a"), ("c/d/e", "c/d"), ("c/d", "c")):
        assert every.index(child) < every.index(parent)
    assert sorted(files) == ["a/x/1.gml", "a/x/2.gml", "a/y.gml", "b/3.gml", "c/d/e/4.gml", "top.gml"]


@pytest.mark.reference_live
def test_directory_walks_equal_reference_live():
    from oracle import refload
    rio = refload.subgraph().io
    with tempfile.TemporaryDirectory() as d:
        _make_tree(d)
        assert sio.get_subdirs(d) == rio.get_subdirs(d)
        assert sio.get_subdirs(d, leaf_only=False) == rio.get_subdirs(d, leaf_only=False)
        assert sio.get_files(d) == rio.get_files(d)
