"""GPU parity of dn4gl_gemm_f32 (3xTF32 tensor-core GEMM with fp32 register accumulation, csrc/gemm3x.cu) against a
float64 product of the same fp32 inputs: both operand layouts, every column-tile width, ragged K / M / N, unaligned
leading dimensions, bias; and the accuracy relative to the fp32 library GEMM it replaces."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 1), (5, 3, 2), (100, 7, 5), (257, 33, 17), (1000, 64, 64), (3000, 130, 200), (15424, 64, 1024),
          (2048, 1152, 64), (777, 96, 129), (4100, 512, 40), (130, 31, 300)]


@pytest.mark.parametrize("N,K,M", SHAPES)
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("with_bias", [False, True])
def test_gemm_matches_float64(device, N, K, M, layout, with_bias):
    from dummynode4graphlearning_b200 import ops
    g = torch.Generator().manual_seed(N * 31 + K * 7 + M + layout)
    a = torch.randn(N, K, generator=g)
    b = torch.randn(M, K, generator=g) if layout == 0 else torch.randn(K, M, generator=g)
    bias = torch.randn(M, generator=g) if with_bias else None
    ref = a.double() @ (b.double().t() if layout == 0 else b.double())
    if with_bias:
        ref = ref + bias.double()
    out = ops.gemm(a.to(device), b.to(device), layout, None if bias is None else bias.to(device)).double().cpu()
    lib32 = (a.to(device) @ (b.to(device).t() if layout == 0 else b.to(device))).double().cpu()
    if with_bias:
        lib32 = lib32 + bias.double()
    # scale: the natural magnitude of an entry, sqrt(K) for unit-variance operands
    scale = max(1.0, float(K) ** 0.5)
    err = float((out - ref).abs().max()) / scale
    err_lib = float((lib32 - ref).abs().max()) / scale
    assert err <= 2e-6, (err, err_lib)
    assert err <= 4 * err_lib + 2e-7, "less accurate than the fp32 library GEMM: %g vs %g" % (err, err_lib)


def test_gemm_large_magnitude_spread_and_views(device):
    """operands with entries over eight orders of magnitude (the hi / lo split must not lose the small ones), and
    the autograd wrappers that route through the kernel (ops.linear, ops.matmul_xw) against float64."""
    from dummynode4graphlearning_b200 import ops
    g = torch.Generator().manual_seed(5)
    a = torch.randn(900, 70, generator=g) * torch.logspace(-4, 4, 70).unsqueeze(0)
    w = torch.randn(48, 70, generator=g)
    ref = a.double() @ w.double().t()
    out = ops.gemm(a.to(device), w.to(device), 0).double().cpu()
    assert float(((out - ref).abs() / (a.double().abs() @ w.double().abs().t())).max()) <= 1e-6

    x = torch.randn(1500, 40, generator=g)      # >= ops.GEMM_MIN_ROWS: the wrappers take the kernel
    W = torch.randn(96, 40, generator=g) * 0.3
    bb = torch.randn(96, generator=g)
    Wk = torch.randn(40, 72, generator=g) * 0.3
    wt1, wt2 = torch.randn(1500, 96, generator=g), torch.randn(1500, 72, generator=g)

    def run(dt, dev, lin, mm):
        xs, Ws, bs, Wks = (t.to(dev, dt).requires_grad_() for t in (x, W, bb, Wk))
        y = (lin(xs, Ws, bs) * wt1.to(dev, dt)).sum() + (mm(xs, Wks) * wt2.to(dev, dt)).sum()
        y.backward()
        return [float(y)] + [t.grad.double().cpu() for t in (xs, Ws, bs, Wks)]

    import torch.nn.functional as F
    assert x.size(0) >= ops.GEMM_MIN_ROWS and ops.GEMM_TENSOR_CORES
    r = run(torch.float64, "cpu", F.linear, lambda u, v: u @ v)
    o = run(torch.float32, device, ops.linear, ops.matmul_xw)
    assert abs(o[0] - r[0]) <= 1e-5 * abs(r[0])
    for got, want in zip(o[1:], r[1:]):
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_counting_model_with_tensor_core_gemm_opt_in(device):
    """the counting models keep their wide products on the library GEMM by default (basemodel._CountingBase.tensor_core_gemm,
    DESIGN.md section 4 K7); with the switch on, RGIN 'small' must give the same prediction and gradients within 1e-5 of the
    default path -- the opt-in path is exercised, and the scope taken in the forward is the one its backward uses."""
    from dummynode4graphlearning_b200 import ops, synth, transforms as T
    from dummynode4graphlearning_b200.graph import BatchedGraph
    from dummynode4graphlearning_b200.subgraph_isomorphism.models import RGIN
    p, g, counts = synth.counting_batch("small", 256, seed=3)
    cfg = dict(synth.counting_config("small"), add_dummy=True)
    mc = T.process_model_config(cfg)
    pd_ = T.sub_add_dummy(T.to_device(p, device), cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"])
    gd_ = T.sub_add_dummy(T.to_device(g, device), cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
    kw = dict({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=64, rep_num_graph_layers=2, rep_num_pattern_layers=2,
              rep_act_func="relu", pred_act_func="relu", pred_net="SumPredictNet", pred_hid_dim=32, emb_net="Equivariant",
              enc_net="Multihot", filter_net="ScalarFilter", pred_with_enc=True, pred_with_deg=True,
              rep_rgin_regularizer="bdd", rep_rgin_num_bases=4)
    results = []
    for flag in (False, True):
        torch.manual_seed(2)
        model = RGIN(**kw)
        with torch.no_grad():
            for n, q in model.named_parameters():
                if "pred_fc2" in n:
                    q.normal_(0.0, 0.1)
        model = model.to(device).train()
        model.tensor_core_gemm = flag
        lib0 = ops.lib().launches
        out = model(BatchedGraph.from_batch(pd_, device), BatchedGraph.from_batch(gd_, device))
        loss = ((out["pred_c"].view(-1) - torch.from_numpy(counts).to(device).float()) ** 2).mean()
        loss.backward()
        results.append((float(loss), [q.grad.detach().clone() for q in model.parameters() if q.grad is not None],
                        ops.lib().launches - lib0))
    (l0, g0, n0), (l1, g1, n1) = results
    assert n1 > n0, "the opt-in path did not issue any additional library call (dn4gl_gemm_f32)"
    assert abs(l1 - l0) <= 1e-5 * abs(l0)
    gmax = max(float(a.abs().max()) for a in g0)
    for a, b in zip(g0, g1):
        assert float((a - b).abs().max()) <= 1e-5 * max(float(a.abs().max()), 1e-3 * gmax)
