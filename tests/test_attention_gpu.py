"""The two attention kernels against an fp32 torch reference: the tcgen05/TMEM encoder kernel (attn_tc.cu, single-pass
softmax with a per-row reference and in-TMEM re-basing) and the mma.sync flash kernel (decoder cross-attention).
Tolerance: relative L2 <= 3e-3 (fp16) / 2e-2 (bf16) — P is rounded to the 16-bit operand type."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, n, T, heads, scale):
    D = qkv.shape[1] // 3
    x = qkv.float().view(n, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax((x[0] @ x[1].transpose(-2, -1)) * scale, -1)
    return (att @ x[2]).transpose(1, 2).reshape(n * T, D)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,T,heads", [(3, 577, 2), (2, 128, 1), (1, 130, 3), (2, 50, 2), (1, 1025, 1)])
def test_attention_matches_reference(cuda_ctx, dtype16, mode, n, T, heads):
    from marie_icr_b200 import ops
    torch.manual_seed(n * T + heads)
    qkv = torch.randn(n * T, 3 * heads * 64, device="cuda").to(dtype16)
    out = ops.attention16(qkv, n, T, 0.125, mode).float()
    ref = _ref(qkv, n, T, heads, 0.125)
    rel = ((out - ref).norm() / ref.norm()).item()
    assert torch.isfinite(out).all()
    assert rel <= (3e-3 if dtype16 == torch.float16 else 2e-2), rel


@pytest.mark.parametrize("mode", [0, 1])
def test_attention_sharp_and_growing_scores(cuda_ctx, dtype16, mode):
    """Scores that grow by far more than the fp16 head-room from key tile to key tile (forces the re-basing path of the
    tcgen05 kernel in every tile) and rows dominated by a single key."""
    from marie_icr_b200 import ops
    torch.manual_seed(7)
    n, T, heads = 2, 577, 2
    D = heads * 64
    qkv = torch.randn(n * T, 3 * D, device="cuda")
    ramp = torch.linspace(0.2, 6.0, T, device="cuda").repeat(n)[:, None]
    qkv[:, D:2 * D] = qkv[:, D:2 * D].abs() * ramp            # keys grow along the sequence
    qkv[:, :D] = qkv[:, :D].abs() * 2.0                       # positive queries: later keys score much higher
    qkv = qkv.to(dtype16)
    out = ops.attention16(qkv, n, T, 0.125, mode).float()
    ref = _ref(qkv, n, T, heads, 0.125)
    rel = ((out - ref).norm() / ref.norm()).item()
    assert torch.isfinite(out).all()
    assert rel <= (4e-3 if dtype16 == torch.float16 else 2e-2), rel


@pytest.mark.parametrize("rows,D", [(577 * 3, 768), (1000, 1024), (37, 128), (5, 768)])
def test_layernorm_matches_torch(cuda_ctx, dtype16, rows, D):
    from marie_icr_b200 import ops
    torch.manual_seed(rows + D)
    x = (torch.randn(rows, D, device="cuda") * 2 + 0.3).to(dtype16)
    g = torch.randn(D, device="cuda") * 0.1 + 1
    b = torch.randn(D, device="cuda") * 0.1
    out = ops.layernorm16(x, g, b, 1e-6).float()
    ref = torch.nn.functional.layer_norm(x.float(), (D,), g, b, 1e-6)
    tol = 2e-3 if dtype16 == torch.float16 else 1.6e-2
    assert (out - ref).abs().max().item() <= tol * ref.abs().max().item()
