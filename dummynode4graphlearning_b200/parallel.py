"""Data-parallel training over mini-batches of independent graphs (one process per GPU).

The reference is single-device; a mini-batch is a block-diagonal union of
independent graphs, so it shards with no data-path exchange.  The only collectives are
  * one all-reduce(sum) of a single flat fp32 gradient bucket per step (0.2 - 0.7 MB for the reference's
    configs: latency-bound, NCCL over NVLink 5), weighted so the result equals the single-process
    mean-loss gradient (rank weight = B_r / B);
  * optionally one all-reduce(max) of the batch-wide padded lengths (Lmax, Lp_max) that enter the counting
    head / label filter through padding (App. A-7, A-14) -- ``sync_padded_lengths``.
BatchNorm statistics stay per rank unless the model was converted with ``torch.nn.SyncBatchNorm``.
Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests/test_parallel_gloo.py).
"""
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(num_items, rank, world_size):
    """contiguous, balanced slice of a batch for `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shard_ranges(work, world_size):
    """contiguous slices [lo, hi) of a mini-batch, one per rank, balanced by per-sample work instead of by count
    (work = nodes + edges of a sample; the reference's BucketSampler, utils/sampler.py:59-64, already
    puts samples of similar size next to each other, so contiguous cuts suffice).  Cut k is placed where the running sum
    of work crosses k / world_size of the total, at the nearer sample boundary; every rank gets at least one sample when
    there are enough samples.  Returns a list of world_size (lo, hi) pairs covering [0, len(work))."""
    n = len(work)
    w = [max(float(x), 0.0) for x in work]
    total = sum(w)
    if world_size <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world_size, 1) - 1)
    prefix = [0.0]
    for x in w:
        prefix.append(prefix[-1] + x)
    cuts, lo = [0], 0
    for k in range(1, world_size):
        target = total * k / world_size
        i = lo
        while i < n and prefix[i + 1] <= target:
            i += 1
        if i < n and (target - prefix[i]) > (prefix[i + 1] - target):   # the boundary after sample i is nearer
            i += 1
        i = max(i, min(lo + 1, n))                 # at least one sample for the previous rank ...
        i = min(i, max(n - (world_size - k), lo))  # ... and leave one for each rank still to come, when possible
        cuts.append(i)
        lo = i
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


class GradientBucket:
    """Flat fp32 view over all trainable parameters' gradients: one collective per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        seen, uniq = set(), []
        for p in self.params:  # shared modules (share_rep_net / share_enc_net) register a tensor twice
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def _ensure(self):
        """built after the first backward: only parameters that actually received a gradient join the bucket
        (EquivariantEmbedding.row_vec never does -- SURVEY.md App. A-13 -- and must stay ``grad is None`` so the
        optimizer keeps skipping it, as in the reference)."""
        if self.flat is None:
            self.active = [p for p in self.params if p.grad is not None]
            if not self.active:
                return
            self.numel = sum(((p.numel() + 3) // 4) * 4 for p in self.active)
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=self.active[0].device)
            off = 0
            self.views = []
            for p in self.active:  # gradients become views into the flat bucket
                n = p.numel()
                view = self.flat[off: off + n].view_as(p)
                view.copy_(p.grad)
                p.grad = view
                self.views.append(view)
                off += ((n + 3) // 4) * 4   # every view 16-byte aligned (vectorised optimizer / copies)

    def zero(self):
        """start of a step.  Gradients are detached from the flat buffer (``p.grad = None``) so that autograd's
        AccumulateGrad nodes adopt the incoming gradient tensors instead of launching one ``grad += new`` kernel per
        parameter; ``gather()`` copies them into the flat buffer with one multi-tensor copy afterwards."""
        for p in self.params:
            p.grad = None

    def gather(self):
        """after backward: every active parameter's gradient ends up in (and ``p.grad`` points into) the flat buffer."""
        self._ensure()
        if self.flat is None:
            return
        src, dst = [], []
        for p, view in zip(self.active, self.views):
            g = p.grad
            if g is view:
                continue
            if g is None:
                view.zero_()                      # did not take part in this step's loss
            else:
                src.append(g if g.dtype == view.dtype else g.to(view.dtype))
                dst.append(view)
            p.grad = view
        if dst:
            torch._foreach_copy_(dst, src)

    def all_reduce(self, local_weight=None):
        """sum over ranks of local_weight * grad (local_weight = B_r / B for mean-reduced losses)."""
        self.gather()
        if self.flat is None:
            return
        if local_weight is not None and local_weight != 1.0:
            self.flat.mul_(local_weight)
        if is_distributed():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def sync_padded_lengths(*lengths):
    """all-reduce(max) of the batch-wide padded lengths so sharded batches reproduce the single-process
    head / filter semantics exactly."""
    if not is_distributed():
        return lengths
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(lengths), dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return tuple(int(x) for x in t.tolist())


def max_over_ranks(value, device=None):
    """max of a python float over ranks (device-timed durations are reported as the max over ranks)."""
    if not is_distributed():
        return value
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
