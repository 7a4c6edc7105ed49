"""CPU: host-side pieces of bench.py -- the sharded CPU-baseline transform equals the single-thread oracle, and the
reference arm prints a contract-shaped JSON line."""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from dummynode4graphlearning_b200 import synth  # noqa: E402


@pytest.mark.parametrize("shape,B,threads", [("proteins", 37, 4), ("mutag", 5, 8), ("proteins", 1, 3)])
def test_sharded_cpu_transform_equals_single_thread(shape, B, threads):
    raw = synth.tu_batch(shape, B, seed=5)
    chunks = bench.split_tu_batch(raw, threads)
    assert sum(c["num_graphs"] for c in chunks) == B and len(chunks) <= min(threads, B)
    assert sum(len(c["src"]) for c in chunks) == len(raw["src"])
    one = bench.cpu_transform(raw, None, 1)
    with ThreadPoolExecutor(threads) as pool:
        many = bench.cpu_transform(raw, pool, threads)
    for a, b in zip(one, many):
        assert np.array_equal(a, b)


def test_reference_arm_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "1", "--graphs", "64"], text=True, cwd=ROOT)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "graphs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and line["higher_is_better"] is True


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                                  text=True, cwd=ROOT, env=env)
    assert out.strip() == ""


@pytest.mark.parametrize("seed", range(12))
def test_tu_conjugate_sizes_match_the_oracle(seed):
    """transforms.tu_conjugate_sizes (host-side size hint that removes the transform's read-back) == the sizes of the
    oracle's edge-to-vertex transform, on the synthetic shapes and on multigraphs with loops / isolated nodes / edgeless
    graphs, with and without the dummy augmentation."""
    import numpy as np
    from dummynode4graphlearning_b200 import synth, transforms as T
    from helpers import nasty_tu_batch
    from oracle import transforms as OT
    rng = np.random.default_rng(seed)
    if seed < 4:
        raw = synth.tu_batch("proteins" if seed % 2 else "mutag", 7 + seed, seed=seed)
    else:
        raw = nasty_tu_batch(rng, int(rng.integers(1, 6)))
    for with_dummy in (True, False):
        ref = OT.tu_conjugate(OT.tu_add_dummy(raw)) if with_dummy else OT.tu_conjugate(raw)
        nodes = np.diff(np.asarray(ref["node_ptr"], np.int64))
        want = (int(ref["node_ptr"][-1]), int(ref["edge_ptr"][-1]), int(nodes.max()))
        assert T.tu_conjugate_sizes(raw, with_dummy) == want, (seed, with_dummy)
