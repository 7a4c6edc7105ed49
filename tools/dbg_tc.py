"""debug probe of the tensor-core backward stage: prints relative errors instead of asserting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dummynode4graphlearning_b200 import ops

dev = torch.device("cuda:0")


def rel(a, ref):
    ref = ref.double()
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


for (K, M) in [(32, 32), (64, 64), (32, 64), (64, 32), (4, 32), (40, 24)]:
    for N in [1, 129, 5000]:
        g = torch.Generator().manual_seed(K * 1000 + M * 10 + N + 1)
        x = torch.randn(N, K, generator=g).to(dev)
        W = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev)
        G = torch.randn(N, M, generator=g).to(dev)
        gx, sp, dW, db = ops.lin_bwd(G, W, x)
        torch.cuda.synchronize()
        print("K=%d M=%d N=%d  gx %.2e  dW %.2e  db %.2e" % (K, M, N, rel(gx, G.double() @ W.double()),
                                                        rel(dW, G.double().t() @ x.double()), rel(db, G.double().sum(0))), flush=True)
