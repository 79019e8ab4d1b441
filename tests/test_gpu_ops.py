"""GPU numerics of the tensor-core building blocks against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 256, 128), (1000, 768, 768), (2356, 3072, 768), (777, 144, 1296),
                                   (5, 16, 16), (129, 48, 432)])
def test_linear_tcgen05(cuda_dev, M, N, K):
    from instageo_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * 0.5).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda_dev).bfloat16()
    b = torch.randn(N, generator=g).to(cuda_dev)
    ref = a.float() @ w.float().t() + b
    scale = max(1.0, ref.abs().max().item())
    out = ops.linear(a, w, b, out_dtype=torch.float32)
    assert (out - ref).abs().max().item() < 1e-3 * scale  # fp32 accumulate: only summation-order noise
    out = ops.linear(a, w, b, act=1)
    assert (out.float() - F.gelu(ref)).abs().max().item() < 1e-2 * scale  # bf16 output rounding (tolerance 2^-8 rel)
    r = torch.randn(M, N, generator=g).to(cuda_dev)
    want = ref + r
    assert (ops.linear(a, w, b, resid=r) - want).abs().max().item() < 1e-3 * scale


@pytest.mark.parametrize("M,N,K", [(1000, 768, 768), (37696, 768, 768), (301, 1024, 4096), (64, 64, 64)])
def test_linear_inplace_residual_reduction(cuda_dev, M, N, K, monkeypatch):
    """x += a @ w.T + b: the bulk-tensor-reduction epilogue (the L2 adds the staged tile into x) gives bit for bit what
    the load / add / store epilogue gives, including the partial last row tile and several tiles per CTA pair."""
    from instageo_b200 import ops
    g = torch.Generator().manual_seed(N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda_dev).bfloat16()
    b = torch.randn(N, generator=g).to(cuda_dev)
    x0 = torch.randn(M, N, generator=g).to(cuda_dev)
    got = ops.linear(a, w, b, resid=x0.clone())
    monkeypatch.setenv("IG_NO_RESID_TMA", "1")
    want = ops.linear(a, w, b, resid=x0.clone())
    monkeypatch.delenv("IG_NO_RESID_TMA")
    assert torch.equal(got, want)
    ref = x0 + a.float() @ w.float().t() + b
    assert (got - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("D", [256, 768, 1024])
def test_layernorm(cuda_dev, D):
    from instageo_b200 import ops
    x = torch.randn(333, D, device=cuda_dev) * 3 + 1
    g, b = torch.randn(D, device=cuda_dev), torch.randn(D, device=cuda_dev)
    ref = F.layer_norm(x, (D,), g, b, 1e-5)
    out = ops.layernorm(x, g, b).float()
    assert (out - ref).abs().max().item() <= 2 ** -8 * max(1.0, ref.abs().max().item()) + 1e-3  # bf16 rounding


@pytest.mark.parametrize("B,N,H", [(1, 128, 1), (2, 197, 4), (2, 589, 12), (3, 64, 2), (1, 1, 1), (2, 129, 16)])
def test_attention(cuda_dev, B, N, H):
    from instageo_b200 import ops
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=cuda_dev).bfloat16()
    out = ops.attention(qkv, B, N, H).float()
    q, k, v = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, D)
    assert (out - ref).abs().max().item() < 1.5e-2  # P and O are rounded to bf16 (tolerance written per north_star)
