"""CPU tests of host-side switches that decide which kernel a product takes (no GPU needed)."""


def test_gemm_scope_and_cpu_routing():
    """ops.gemm_tensor_cores scopes nest and restore; CPU tensors never take the tensor-core GEMM (the wrappers fall through
    to the library on the host, where the oracle and the gloo tests run), and the GIN-MLP dispatcher applies a module as it
    is when no CUDA path applies."""
    import torch
    import torch.nn as nn
    from dummynode4graphlearning_b200 import ops
    assert ops._gemm_scope[0] is None
    with ops.gemm_tensor_cores(False):
        assert ops._gemm_scope[0] is False
        with ops.gemm_tensor_cores(True):
            assert ops._gemm_scope[0] is True
        with ops.gemm_tensor_cores(None):      # None: keep the enclosing choice
            assert ops._gemm_scope[0] is False
        assert ops._gemm_scope[0] is False
    assert ops._gemm_scope[0] is None
    x, w, b = torch.randn(2000, 12), torch.randn(7, 12), torch.randn(7)
    assert not ops._use_gemm(x, w)
    xr = x.clone().requires_grad_()
    y = ops.linear(xr, w, b)
    assert torch.allclose(y, torch.nn.functional.linear(x, w, b), atol=1e-6)
    y.sum().backward()
    assert torch.allclose(xr.grad, torch.ones(2000, 7) @ w, atol=1e-5)
    seq = nn.Sequential(ops.Linear(12, 96), nn.BatchNorm1d(96), nn.ReLU(), ops.Linear(96, 96), nn.BatchNorm1d(96), nn.ReLU())
    assert ops.gin_mlp_wide_ok(seq) and not ops.gin_mlp_fusable(seq)
    torch.manual_seed(0)
    ref = seq(x)
    seq2 = nn.Sequential(ops.Linear(12, 96), nn.BatchNorm1d(96), nn.ReLU(), ops.Linear(96, 96), nn.BatchNorm1d(96), nn.ReLU())
    seq2.load_state_dict(seq.state_dict())
    seq2[1].running_mean.zero_(); seq2[1].running_var.fill_(1.0); seq2[4].running_mean.zero_(); seq2[4].running_var.fill_(1.0)
    assert torch.allclose(ops.apply_gin_mlp(seq2, x), ref, atol=1e-6)      # CPU rows: the module as it is
    seq.eval()
    assert not ops.gin_mlp_wide_ok(seq)
