"""tensor-core MLP stages (csrc/mlp_tc.cu) through the C ABI against float64 torch references.

Tolerance: north_star's 1e-5 relative (max-abs error / max-abs reference) for fp32 work; the 3xTF32 GEMMs are
expected around 1e-6."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

TOL = 1e-5


def rel(a, ref):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("K,M", [(32, 32), (64, 64), (32, 64), (64, 32), (2, 32), (7, 32), (32, 2), (40, 24)])
@pytest.mark.parametrize("N", [1, 127, 128, 129, 5000, 70001])
def test_lin_fwd_plain(K, M, N):
    from dummynode4graphlearning_b200 import ops
    g = torch.Generator().manual_seed(K * 1000 + M * 10 + N)
    x = torch.randn(N, K, generator=g).to(dev())
    W = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev())
    b = torch.randn(M, generator=g).to(dev())
    y, rec = ops.lin_fwd(x, W, b)
    assert rec is None
    ref = x.double() @ W.double().t() + b.double()
    assert rel(y, ref) <= TOL, rel(y, ref)
    y2, _ = ops.lin_fwd(x, W, None)
    assert rel(y2, x.double() @ W.double().t()) <= TOL


@pytest.mark.parametrize("K,M", [(32, 32), (64, 64), (8, 32), (32, 64)])
@pytest.mark.parametrize("N", [300, 20000])
def test_lin_fwd_bn_stats_and_prologue(K, M, N):
    from dummynode4graphlearning_b200 import ops
    g = torch.Generator().manual_seed(7 + K + M + N)
    x = (torch.randn(N, K, generator=g) * 2 + 3).to(dev())
    W = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev())
    b = torch.randn(M, generator=g).to(dev())
    gamma, beta = (torch.rand(M, generator=g) + 0.5).to(dev()), torch.randn(M, generator=g).to(dev())
    rm, rv = torch.zeros(M, device=dev()), torch.ones(M, device=dev())
    nbt = torch.zeros((), dtype=torch.int64, device=dev())
    bn = dict(gamma=gamma, beta=beta, eps=1e-5, momentum=0.1, running_mean=rm, running_var=rv, num_batches_tracked=nbt)
    y, rec = ops.lin_fwd(x, W, b, bn=bn)
    yd = x.double() @ W.double().t() + b.double()
    mean, var = yd.mean(0), yd.var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    assert rel(rec[:M], mean) <= TOL
    assert rel(rec[M:2 * M], rstd) <= TOL
    assert rel(rec[2 * M:3 * M], gamma.double() * rstd) <= TOL
    assert torch.equal(rec[3 * M:], beta)
    assert rel(rm, 0.1 * mean) <= TOL
    assert rel(rv, 0.9 + 0.1 * yd.var(0, unbiased=True)) <= TOL
    assert int(nbt) == 1
    # second stage: BN + ReLU of the first folded into the prologue
    W2 = (torch.randn(M, M, generator=g) / M ** 0.5).to(dev())
    y2, _ = ops.lin_fwd(y, W2, None, in_bn=rec, in_act=ops.ACT_RELU)
    a = torch.relu((yd - mean) * rstd * gamma.double() + beta.double())
    assert rel(y2, a @ W2.double().t()) <= TOL
    h = ops.bn_act(y, rec, ops.ACT_RELU)
    assert rel(h, a) <= TOL
    # repeated launch is bit-identical (fixed-order reductions, counter left at zero)
    bn2 = dict(gamma=gamma, beta=beta, eps=1e-5, momentum=0.1)
    _, rec_b = ops.lin_fwd(x, W, b, bn=bn2)
    _, rec_c = ops.lin_fwd(x, W, b, bn=bn2)
    assert torch.equal(rec_b, rec_c) and torch.equal(rec_b[:3 * M], rec[:3 * M])


def test_bn_stats_degenerate_channel():
    """a channel that is constant up to rounding noise and far from zero (the C2 CONJ features produce these)."""
    from dummynode4graphlearning_b200 import ops
    N, K, M = 50000, 32, 32
    g = torch.Generator().manual_seed(3)
    x = torch.zeros(N, K)
    x[:, 0] = 1.0
    x[:, 1] = (torch.rand(N, generator=g) < 0.001).float()
    W = torch.randn(M, K, generator=g)
    b = torch.randn(M, generator=g) * 10
    x, W, b = x.to(dev()), W.to(dev()), b.to(dev())
    y, rec = ops.lin_fwd(x, W, b, bn=dict(gamma=None, beta=None, eps=1e-5, momentum=0.1))
    yd = x.double() @ W.double().t() + b.double()
    assert rel(rec[:M], yd.mean(0)) <= 1e-6
    var = yd.var(0, unbiased=False)
    assert float(((1.0 / rec[M:2 * M].double() ** 2 - 1e-5) - var).abs().max() / var.max()) <= 1e-4


@pytest.mark.parametrize("K,M", [(32, 32), (64, 64), (32, 64), (64, 32), (4, 32), (40, 24)])
@pytest.mark.parametrize("N", [1, 129, 5000, 70001])
def test_lin_bwd_plain(K, M, N):
    from dummynode4graphlearning_b200 import ops
    g = torch.Generator().manual_seed(K * 1000 + M * 10 + N + 1)
    x = torch.randn(N, K, generator=g).to(dev())
    W = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev())
    G = torch.randn(N, M, generator=g).to(dev())
    gx, sp, dW, db = ops.lin_bwd(G, W, x)
    assert sp is None
    assert rel(gx, G.double() @ W.double()) <= TOL, rel(gx, G.double() @ W.double())
    assert rel(dW, G.double().t() @ x.double()) <= TOL, rel(dW, G.double().t() @ x.double())
    assert rel(db, G.double().sum(0)) <= TOL
    gx2, _, dW2, _ = ops.lin_bwd(G, W, x, want_gx=False)
    assert gx2 is None and torch.equal(dW2, dW)
    # activation prologue without BN (Linear, act, Linear MLPs): X' = act(X), GX masked by act'
    for act, slope in ((ops.ACT_RELU, 0.0), (ops.ACT_LEAKY_RELU, 1 / 5.5)):
        gx, _, dW, _ = ops.lin_bwd(G, W, x, in_act=act, in_slope=slope)
        xd = x.double()
        xa = torch.where(xd > 0, xd, slope * xd)
        da = torch.where(xd > 0, torch.ones_like(xd), torch.full_like(xd, slope))
        assert rel(gx, (G.double() @ W.double()) * da) <= TOL
        assert rel(dW, G.double().t() @ xa) <= TOL


class _RefMlp(nn.Module):
    def __init__(self, din, d):
        super().__init__()
        self.seq = nn.Sequential(nn.Linear(din, d), nn.BatchNorm1d(d), nn.ReLU(), nn.Linear(d, d), nn.BatchNorm1d(d),
                                 nn.ReLU())


def _ref_forward_with_masks(seq, z, mask1, mask2):
    """the float64 module evaluated with GIVEN ReLU masks: with ~1e6 pre-activations a few lie within fp32 rounding of
    zero, where the fp32 and fp64 evaluations legitimately take different sides of the kink (one row of the gradient
    changes by O(1)); fixing the masks to the ones the device used compares everything else at full precision."""
    l1, n1, _, l2, n2, _ = seq
    return n2(l2(n1(l1(z)) * mask1)) * mask2


@pytest.mark.parametrize("din,d", [(32, 32), (64, 64), (4, 32), (32, 64)])
@pytest.mark.parametrize("N", [700, 30000])
def test_gin_mlp_chain_against_torch_float64(din, d, N):
    from dummynode4graphlearning_b200 import ops
    torch.manual_seed(din + d + N)
    ref = _RefMlp(din, d).double()
    with torch.no_grad():
        for n in (ref.seq[1], ref.seq[4]):
            n.weight.uniform_(0.5, 1.5)
            n.bias.normal_()
    mine = _RefMlp(din, d)
    mine.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in ref.state_dict().items()})
    mine = mine.to(dev())
    z = torch.randn(N, din)
    gh = torch.randn(N, d)
    zm = z.to(dev()).requires_grad_(True)
    assert ops.gin_mlp_fusable(mine.seq)
    hm = ops.gin_mlp(mine.seq, zm)
    hm.backward(gh.to(dev()))
    with torch.no_grad():   # the masks the device used
        l1, n1 = mine.seq[0], mine.seq[1]
        y1, rec1 = ops.lin_fwd(zm.detach(), l1.weight, l1.bias, bn=dict(gamma=n1.weight, beta=n1.bias, eps=n1.eps, momentum=0.1))
        mask1 = (ops.bn_act(y1, rec1, ops.ACT_RELU) > 0).double().cpu()
        mask2 = (hm > 0).double().cpu()
    zr = z.double().requires_grad_(True)
    hr = _ref_forward_with_masks(ref.seq, zr, mask1, mask2)
    hr.backward(gh.double())
    assert rel(hm, hr) <= TOL, rel(hm, hr)
    assert rel(zm.grad, zr.grad) <= TOL, rel(zm.grad, zr.grad)
    for (k, pm), (_, pr) in zip(mine.named_parameters(), ref.named_parameters()):
        # biases of a Linear followed by BatchNorm have an exactly-zero gradient: compare absolutely
        if k in ("seq.0.bias", "seq.3.bias"):
            assert float(pm.grad.abs().max()) <= 1e-4 * float(gh.abs().sum() / N), k
        else:
            assert rel(pm.grad, pr.grad) <= TOL, (k, rel(pm.grad, pr.grad))
    for k in ("running_mean", "running_var"):
        for i in (1, 4):
            assert rel(getattr(mine.seq[i], k), getattr(ref.seq[i], k)) <= TOL
    assert int(mine.seq[1].num_batches_tracked) == 1


@pytest.mark.parametrize("M", [32, 64, 24])
@pytest.mark.parametrize("mean", [False, True])
def test_bn_act_pool_and_segment_ids(M, mean):
    from dummynode4graphlearning_b200 import ops
    g = torch.Generator().manual_seed(M)
    sizes = torch.randint(1, 300, (57,), generator=g)
    sizes[3] = 2500
    N = int(sizes.sum())
    seg = torch.zeros(58, dtype=torch.int32)
    seg[1:] = torch.cumsum(sizes, 0)
    Y = torch.randn(N, M, generator=g).to(dev())
    rec = torch.cat([torch.randn(M, generator=g) * 0.1, torch.rand(M, generator=g) + 0.5, torch.rand(M, generator=g) + 0.5,
                     torch.randn(M, generator=g) * 0.1]).to(dev())
    h, pooled = ops.bn_act_pool(Y, rec, seg.to(dev()), mean=mean)
    href = torch.relu((Y.double() - rec[:M].double()) * rec[2 * M:3 * M].double() + rec[3 * M:].double())
    ids = torch.repeat_interleave(torch.arange(57), sizes)
    assert torch.equal(ops.segment_ids(seg.to(dev()), N).cpu(), ids.to(torch.int32))
    pref = torch.zeros(57, M, dtype=torch.float64).index_add_(0, ids, href.cpu())
    if mean:
        pref = pref / sizes.double().unsqueeze(1)
    assert rel(h, href) <= 1e-6
    assert rel(pooled, pref) <= TOL
    assert torch.equal(h, ops.bn_act(Y, rec))


@pytest.mark.parametrize("K,M", [(32, 32), (64, 64), (64, 32)])
def test_lin_bwd_with_readout_gradient(K, M):
    """G + gseg[row2seg] folded into the prologue == the materialised sum (and G = None == broadcast only)."""
    from dummynode4graphlearning_b200 import ops
    g = torch.Generator().manual_seed(K + M)
    N, B = 9000, 40
    ids = torch.sort(torch.randint(0, B, (N,), generator=g)).values.to(torch.int32).to(dev())
    x = torch.randn(N, K, generator=g).to(dev())
    W = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev())
    G = torch.randn(N, M, generator=g).to(dev())
    gseg = torch.randn(B, M, generator=g).to(dev())
    Y = torch.randn(N, M, generator=g).to(dev())
    rec = torch.cat([torch.zeros(M), torch.ones(M), torch.rand(M, generator=g) + 0.5, torch.zeros(M)]).to(dev())
    for Gin in (G, None):
        tot = gseg[ids.long()] + (G if Gin is not None else 0)
        s_a = ops.bn_bwd_sums(tot, Y, rec)
        s_b = ops.bn_bwd_sums(Gin, Y, rec, gseg=gseg, row2seg=ids)
        assert rel(s_b, s_a) <= 1e-6
        ra = ops.lin_bwd(tot, W, x, Yout=Y, bn=rec, sums=s_a)
        rb = ops.lin_bwd(Gin, W, x, Yout=Y, bn=rec, sums=s_a, gseg=gseg, row2seg=ids)
        for a, b in zip(ra, rb):
            if a is not None:
                assert rel(b, a) <= 1e-6


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "none"])
@pytest.mark.parametrize("d,N", [(64, 20000), (32, 333)])
def test_mlp2_against_torch_float64(act, d, N):
    """Linear, act, Linear of the counting models (rgin.py:52, dmpnn.py:47,55)."""
    from dummynode4graphlearning_b200 import ops
    from dummynode4graphlearning_b200.subgraph_isomorphism.utils import map_activation_str_to_layer
    torch.manual_seed(d + N)
    seq = nn.Sequential(nn.Linear(d, d), map_activation_str_to_layer(act), nn.Linear(d, d))
    ref = nn.Sequential(nn.Linear(d, d), map_activation_str_to_layer(act), nn.Linear(d, d)).double()
    ref.load_state_dict({k: v.double() for k, v in seq.state_dict().items()})
    seq = seq.to(dev())
    assert ops.mlp2_fusable(seq, force=True) and ops.mlp2_fusable(seq) == ops.MLP2_TENSOR_CORES   # default on since round 2
    x = torch.randn(N, d)
    g = torch.randn(N, d)
    xm = x.to(dev()).requires_grad_(True)
    ym = ops.mlp2(seq, xm)
    ym.backward(g.to(dev()))
    xr = x.double().requires_grad_(True)
    # same side of the activation kink as the device took (see _ref_forward_with_masks)
    with torch.no_grad():
        y1 = ops.lin_fwd(xm.detach(), seq[0].weight, seq[0].bias)[0].cpu()
    pre = ref[0](xr)
    if act == "none":
        a = pre
    else:
        slope = 0.0 if act == "relu" else 1 / 5.5
        pos = (y1 > 0).double()
        a = pre * (pos + (1 - pos) * slope)
    yr = ref[2](a)
    yr.backward(g.double())
    assert rel(ym, yr) <= TOL
    assert rel(xm.grad, xr.grad) <= TOL
    for (k, pm), (_, pr) in zip(seq.named_parameters(), ref.named_parameters()):
        assert rel(pm.grad, pr.grad) <= TOL, k
