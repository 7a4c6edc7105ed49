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
        self._peer = None       # PeerAllReduce once set up, False if unavailable

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
        """sum over ranks of local_weight * grad (local_weight = B_r / B for mean-reduced losses).  On GPUs of one node:
        ONE kernel per rank that reads every rank's bucket over NVLink peer memory (``PeerAllReduce``); otherwise -- CPU /
        gloo, more than 8 ranks, peer mappings unavailable -- a scale kernel and the backend's all-reduce."""
        self.gather()
        if self.flat is None:
            return
        if is_distributed():
            if self._peer is None and self.flat.is_cuda and not torch.cuda.is_current_stream_capturing():
                self._peer = PeerAllReduce.create(self.flat)          # collective: every rank reaches this at its first step
            if self._peer:
                self._peer.run(1.0 if local_weight is None else float(local_weight))
                return
        if local_weight is not None and local_weight != 1.0:
            self.flat.mul_(local_weight)
        if is_distributed():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


class PeerAllReduce:
    """the flat gradient bucket summed over the ranks of ONE node by ``dn4gl_peer_allreduce_f32`` (csrc/peer.cu): every
    rank exposes ``weight * bucket`` in a buffer that every other rank has mapped (CUDA IPC handles exchanged once through
    the process group) and adds all ranks' exposed buffers in rank order inside one kernel -- no library collective and no
    separate scaling kernel on the step's critical path.

    ``create`` is itself a collective and returns False (on every rank alike) when any rank cannot set the mappings up:
    backend not NCCL, more than 8 ranks, ``DN4GL_PEER_ALLREDUCE=0``, no peer access between two of the devices, IPC
    refused by the platform; the caller then stays on ``dist.all_reduce``."""

    MAX_WORLD = 8
    SIGNAL_WORDS = 2048     # DN4GL_PEER_SIGNAL_WORDS

    def __init__(self, flat, area, exposed_ptrs, signal_ptrs, mapped, rank, world):
        import ctypes
        self.flat, self.area, self.rank, self.world = flat, area, rank, world
        self._mapped = mapped                                      # base pointers of the peer mappings (dn4gl_ipc_close)
        self._exposed = (ctypes.c_void_p * world)(*exposed_ptrs)
        self._signals = (ctypes.c_void_p * world)(*signal_ptrs)

    @staticmethod
    def _export(t):
        """(cudaIpcMemHandle bytes of the allocation holding t, byte offset of t inside it) -- torch's own export of the
        caching allocator's segment (storage._share_cuda_)"""
        h = t.untyped_storage()._share_cuda_()
        handle = bytes(h[1])
        if len(handle) == 66:       # torch >= 2.5: version byte + kind byte ('c' = cudaMalloc segment) + cudaIpcMemHandle_t
            if handle[1:2] != b"c":
                raise RuntimeError("the buffer lives in an expandable segment: no cudaIpcMemHandle for it")
            handle = handle[2:]
        if len(handle) != 64:
            raise RuntimeError("unexpected CUDA IPC handle of %d bytes" % len(handle))
        return handle, int(h[3]) + t.storage_offset() * t.element_size()

    @classmethod
    def create(cls, flat):
        import ctypes
        import os
        from ._lib import lib
        world, rank = dist.get_world_size(), dist.get_rank()
        n = flat.numel()
        ok, err, mine, area = True, "", None, None
        try:
            if os.environ.get("DN4GL_PEER_ALLREDUCE", "1") == "0" or dist.get_backend() != "nccl" or world > cls.MAX_WORLD \
                    or n % 4 != 0 or flat.dtype != torch.float32:
                raise RuntimeError("not applicable")
            # ONE allocation per rank: [signal words | exposed buffer 0 | exposed buffer 1], zeroed once
            area = torch.zeros(cls.SIGNAL_WORDS + 2 * n, dtype=torch.float32, device=flat.device)
            torch.cuda.synchronize(flat.device)
            mine = cls._export(area)
        except Exception as e:      # noqa: BLE001 -- any failure here means "use the library collective", on every rank
            ok, err = False, repr(e)
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        bases, mapped = [], []
        if ok and all(h is not None for h in handles):
            try:
                with torch.cuda.device(flat.device):
                    for r, h in enumerate(handles):
                        if r == rank:
                            bases.append(area.data_ptr())
                            continue
                        base = ctypes.c_void_p()
                        lib().call("dn4gl_ipc_open", h[0], ctypes.byref(base))
                        mapped.append(base.value)
                        bases.append(base.value + h[1])
            except Exception as e:      # noqa: BLE001
                ok, err = False, repr(e)
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=flat.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            for base in mapped:
                lib().call("dn4gl_ipc_close", base)
            if err and "not applicable" not in err:
                import warnings
                warnings.warn("peer-memory all-reduce unavailable on rank %d (%s): using dist.all_reduce" % (rank, err))
            return False
        return cls(flat, area, [b + 4 * cls.SIGNAL_WORDS for b in bases], bases, mapped, rank, world)

    def run(self, weight):
        from ._lib import lib
        from .graph import _stream
        lib().call("dn4gl_peer_allreduce_f32", self.flat.data_ptr(), self.flat.numel(), float(weight), self._exposed,
                   self._signals, self.rank, self.world, _stream())


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
