"""debug probe of the fused GIN-MLP backward chain: per-piece errors against float64, per-tile error pattern."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dummynode4graphlearning_b200 import ops

dev = torch.device("cuda:0")


def rel(a, ref):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def tile_pattern(a, ref, name):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    err = (a - ref).abs().max(dim=1).values / ref.abs().max()
    nt = (err.numel() + 127) // 128
    bad = [i for i in range(nt) if float(err[i * 128:(i + 1) * 128].max()) > 1e-4]
    print("   %s: %d bad tiles of %d: %s" % (name, len(bad), nt, bad[:40]))


for (din, d, N) in [(64, 64, 30000), (32, 32, 70000), (64, 64, 700)]:
    torch.manual_seed(1)
    z = torch.randn(N, din)
    W1, b1 = torch.randn(d, din) / din ** 0.5, torch.randn(d)
    W2, b2 = torch.randn(d, d) / d ** 0.5, torch.randn(d)
    g1, be1, g2, be2 = torch.rand(d) + 0.5, torch.randn(d), torch.rand(d) + 0.5, torch.randn(d)
    gh = torch.randn(N, d)
    D = lambda t: t.double()
    # float64 reference, piece by piece
    y1 = D(z) @ D(W1).t() + D(b1)
    m1, v1 = y1.mean(0), y1.var(0, unbiased=False); r1 = 1 / torch.sqrt(v1 + 1e-5)
    xh1 = (y1 - m1) * r1; p1 = xh1 * D(g1) + D(be1); a1 = torch.relu(p1)
    y2 = a1 @ D(W2).t() + D(b2)
    m2, v2 = y2.mean(0), y2.var(0, unbiased=False); r2 = 1 / torch.sqrt(v2 + 1e-5)
    xh2 = (y2 - m2) * r2; p2 = xh2 * D(g2) + D(be2)
    gm2 = D(gh) * (p2 > 0)
    s21, s22 = gm2.sum(0), (gm2 * xh2).sum(0)
    gy2 = D(g2) * r2 * (gm2 - s21 / N - xh2 * s22 / N)
    ga1 = (gy2 @ D(W2)) * (p1 > 0)
    s11, s12 = ga1.sum(0), (ga1 * xh1).sum(0)
    gy1 = D(g1) * r1 * (ga1 - s11 / N - xh1 * s12 / N)
    gz = gy1 @ D(W1)
    # device
    c = lambda t: t.to(dev)
    Y1, rec1 = ops.lin_fwd(c(z), c(W1), c(b1), bn=dict(gamma=c(g1), beta=c(be1), eps=1e-5, momentum=0.1))
    Y2, rec2 = ops.lin_fwd(Y1, c(W2), c(b2), in_bn=rec1, in_act=ops.ACT_RELU, bn=dict(gamma=c(g2), beta=c(be2), eps=1e-5, momentum=0.1))
    print("din=%d d=%d N=%d: y1 %.1e y2 %.1e" % (din, d, N, rel(Y1, y1), rel(Y2, y2)))
    sums2 = ops.bn_bwd_sums(c(gh), Y2, rec2, ops.ACT_RELU)
    print("   sums2 %.1e %.1e" % (rel(sums2[:d], s21), rel(sums2[d:], s22)))
    GA1, sums1, dW2, db2 = ops.lin_bwd(c(gh), c(W2), Y1, Yout=Y2, bn=rec2, sums=sums2, g_masked=False, in_bn=rec1, in_act=ops.ACT_RELU)
    print("   ga1 %.1e sums1 %.1e %.1e dW2 %.1e db2(abs) %.1e" % (rel(GA1, ga1), rel(sums1[:d], s11), rel(sums1[d:], s12), rel(dW2, gy2.t() @ a1), float(db2.abs().max())))
    tile_pattern(GA1, ga1, "ga1")
    GZ, _, dW1, db1 = ops.lin_bwd(GA1, c(W1), c(z), Yout=Y1, bn=rec1, sums=sums1, g_masked=True)
    print("   gz %.1e dW1 %.1e" % (rel(GZ, gz), rel(dW1, gy1.t() @ D(z))))
    tile_pattern(GZ, gz, "gz")
    # the same stage fed with exact inputs
    GZ2, _, dW1b, _ = ops.lin_bwd(c(ga1.float()), c(W1), c(z), Yout=Y1, bn=rec1, sums=c(torch.cat([s11, s12]).float()), g_masked=True)
    print("   gz(exact inputs) %.1e dW1 %.1e" % (rel(GZ2, gz), rel(dW1b, gy1.t() @ D(z))))
    tile_pattern(GZ2, gz, "gz(exact)")
    torch.cuda.synchronize()
