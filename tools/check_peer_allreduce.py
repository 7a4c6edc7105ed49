#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_peer_allreduce.py: dn4gl_peer_allreduce_f32 against dist.all_reduce on the same
random buckets (several sizes, several rounds so that epochs and the two exposed buffers alternate), bit-equality of the
result across ranks, and the time per call of both (CUDA events, 200 calls each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dummynode4graphlearning_b200.parallel import PeerAllReduce


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    worst = 0.0
    for n in (4, 4 * 600, 4 * 2401, 4 * 75001, 4 * 262144):
        torch.manual_seed(1000 * n % 7919 + rank)
        ref = torch.randn(n, device=dev)
        flat = torch.empty_like(ref)
        peer = PeerAllReduce.create(flat)
        assert peer, "peer-memory all-reduce could not be set up"
        w = 0.125 + 0.25 * rank
        for it in range(6):
            flat.copy_(ref * (it + 1))
            peer.run(w)
            expect = ref * (it + 1) * w
            dist.all_reduce(expect)
            err = float((flat - expect).abs().max() / expect.abs().max())
            worst = max(worst, err)
            gathered = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            same = all(torch.equal(gathered[0], g) for g in gathered)
            assert err <= 2e-6 and same, (n, it, err, same)
        # timing
        def timed(fn, k=200):
            for _ in range(20):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                fn()
            e1.record(); torch.cuda.synchronize()
            return 1e3 * e0.elapsed_time(e1) / k
        t_peer = timed(lambda: peer.run(w))
        t_nccl = timed(lambda: (flat.mul_(w), dist.all_reduce(flat)))
        if rank == 0:
            print("n=%d floats: peer %.2f us, scale + nccl %.2f us per call (back to back, world %d)" % (n, t_peer, t_nccl, world), flush=True)
    if rank == 0:
        print("ok: worst relative difference to dist.all_reduce %.2e, identical bits on every rank" % worst, flush=True)
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
