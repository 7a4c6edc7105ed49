"""Host-side data-parallel logic on CPU: world_size-2 ``gloo`` process groups (the N>1 path of bench.py / pipelines.py).

Checks that sharding a mini-batch of independent graphs over ranks and all-reducing ONE flat gradient bucket with
rank weights B_r / B reproduces the single-process mean-loss gradient, that parameters which never
receive a gradient stay ``grad is None`` (App. A-13), and the two tiny auxiliary collectives (padded lengths, timing).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dummynode4graphlearning_b200.parallel import (GradientBucket, balanced_shard_ranges, max_over_ranks, shard_range,
                                                   sync_padded_lengths)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Toy(torch.nn.Module):
    """per-graph readout + head with a parameter that never gets a gradient (like EquivariantEmbedding.row_vec)."""

    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(6, 5)
        self.head = torch.nn.Linear(5, 1)
        self.unused = torch.nn.Parameter(torch.ones(3))

    def forward(self, x, seg):   # x (N, 6), seg (N,) graph index -> (B, 1)
        h = torch.relu(self.lin(x))
        B = int(seg.max()) + 1
        pooled = torch.zeros(B, 5).index_add_(0, seg, h)
        return self.head(pooled)


def _make_batch(num_graphs, seed=0):
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(2, 9, (num_graphs,), generator=g)
    x = torch.randn(int(sizes.sum()), 6, generator=g)
    y = torch.randn(num_graphs, 1, generator=g)
    return sizes, x, y


def _slice(sizes, x, y, lo, hi):
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
    xs = x[ptr[lo]:ptr[hi]]
    seg = torch.repeat_interleave(torch.arange(hi - lo), sizes[lo:hi])
    return xs, seg, y[lo:hi]


def _worker(rank, world, port, num_graphs, out_dir, by_work=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = _Toy()
        bucket = GradientBucket(model.parameters())
        sizes, x, y = _make_batch(num_graphs)
        lo, hi = shard_range(num_graphs, rank, world)
        if by_work:   # contiguous cuts balanced by graph size instead of by count
            lo, hi = balanced_shard_ranges(sizes.tolist(), world)[rank]
        for step in range(2):   # second step exercises the flat-bucket path (zero() on views)
            bucket.zero()
            xs, seg, ys = _slice(sizes, x, y, lo, hi)
            loss = torch.nn.functional.mse_loss(model(xs, seg), ys)   # mean over the LOCAL graphs
            loss.backward()
            bucket.all_reduce((hi - lo) / num_graphs)
        assert model.unused.grad is None
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        lens = sync_padded_lengths(int(sizes[lo:hi].max()), 3 + rank)
        slowest = max_over_ranks(1.0 + rank)
        if rank == 0:
            torch.save({"grads": grads, "lens": lens, "slowest": slowest}, os.path.join(out_dir, "dp.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_graphs,by_work", [(8, False), (7, False), (9, True)])   # even, ragged, work-balanced shards
def test_sharded_gradient_equals_single_process(tmp_path, num_graphs, by_work):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), num_graphs, str(tmp_path), by_work), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    torch.manual_seed(0)
    model = _Toy()
    sizes, x, y = _make_batch(num_graphs)
    xs, seg, ys = _slice(sizes, x, y, 0, num_graphs)
    torch.nn.functional.mse_loss(model(xs, seg), ys).backward()
    for n, p in model.named_parameters():
        if p.grad is None:
            assert n not in got["grads"]
        else:
            torch.testing.assert_close(got["grads"][n], p.grad, rtol=1e-5, atol=1e-6)
    assert got["lens"] == (int(sizes.max()), 3 + world - 1)
    assert got["slowest"] == float(world)


def test_shard_range_partitions():
    for n in (0, 1, 5, 512, 1113):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            lens = [b - a for a, b in spans]
            assert max(lens) - min(lens) <= 1


def test_balanced_shard_ranges():
    """contiguous cuts by work: cover the batch, never empty while samples remain, and no rank carries more than the ideal
    share plus one sample; with size-sorted batches (BucketSampler) they beat the by-count split."""
    import numpy as np
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 64, 512):
        for w in (1, 2, 3, 8):
            work = rng.integers(10, 600, n).astype(float)
            spans = balanced_shard_ranges(work, w)
            assert len(spans) == w and spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            if n >= w:
                assert all(b > a for a, b in spans)
                loads = [work[a:b].sum() for a, b in spans]
                assert max(loads) <= work.sum() / w + work.max() + 1e-9
    work = np.sort(rng.integers(10, 2000, 512)).astype(float)          # one size-sorted batch of 512 samples
    by_work = max(work[a:b].sum() for a, b in balanced_shard_ranges(work, 8))
    by_count = max(work[a:b].sum() for a, b in (shard_range(512, r, 8) for r in range(8)))
    assert by_work < 0.65 * by_count
    assert balanced_shard_ranges([], 4) == [(0, 0)] * 4 and balanced_shard_ranges([5.0, 1.0], 1) == [(0, 2)]
    assert balanced_shard_ranges([0.0, 0.0, 0.0], 3) == [(0, 1), (1, 2), (2, 3)]


def test_bucket_single_process_is_identity():
    torch.manual_seed(1)
    model = _Toy()
    bucket = GradientBucket(model.parameters())
    sizes, x, y = _make_batch(5)
    xs, seg, ys = _slice(sizes, x, y, 0, 5)
    torch.nn.functional.mse_loss(model(xs, seg), ys).backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    bucket.all_reduce(1.0)
    for n, p in model.named_parameters():
        if n in ref:
            assert p.grad.data_ptr() >= bucket.flat.data_ptr()   # gradients are views into the flat bucket
            torch.testing.assert_close(p.grad, ref[n])
    assert model.unused.grad is None
