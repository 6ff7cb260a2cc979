"""tap-GEMM (tcgen05/TMA) against torch fp32 references on the same 16-bit operands (fp16 and bf16).

Tolerance: operands are identical bf16 values, accumulation is fp32 on both sides, so the only differences
are summation order and the final bf16 rounding of the output: |err| <= 2^-8 * |ref| + 1e-3*scale.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(out, ref, name):
    out = out.float()
    scale = ref.abs().max().item() + 1e-6
    err = (out - ref).abs().max().item()
    rel = err / scale
    assert rel < 1e-2, f"{name}: max abs err {err} (scale {scale}, rel {rel})"
    return rel


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 128), (577 * 3, 768, 768), (1000, 2304, 768),
                                   (37, 1024, 4096), (4096, 50265, 1024), (130, 96, 192)])
def test_gemm_plain(cuda_ctx, dtype16, M, N, K):
    from marie_icr_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(dtype16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dtype16)
    out = ops.gemm16(a, w)
    ref = a.float() @ w.float().t()
    _close(out, ref, f"gemm {M}x{N}x{K}")


def test_gemm_epilogues(cuda_ctx, dtype16):
    from marie_icr_b200 import ops
    torch.manual_seed(1)
    M, N, K = 700, 768, 256
    a = torch.randn(M, K, device="cuda").to(dtype16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dtype16)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(dtype16)
    base = a.float() @ w.float().t() + bias
    _close(ops.gemm16(a, w, bias=bias), base, "bias")
    _close(ops.gemm16(a, w, bias=bias, act=ops.ACT_RELU), base.relu(), "relu")
    _close(ops.gemm16(a, w, bias=bias, act=ops.ACT_GELU), F.gelu(base), "gelu")
    _close(ops.gemm16(a, w, bias=bias, residual=res), base + res.float(), "residual")
    out32 = ops.gemm16(a, w, bias=bias, out_dtype=torch.float32)
    assert out32.dtype == torch.float32
    assert (out32 - base).abs().max().item() < 2e-3 * base.abs().max().item()


@pytest.mark.parametrize("n,h,w,cin,cout,dil", [(1, 16, 128, 64, 64, 1), (2, 24, 200, 128, 256, 1),
                                               (1, 20, 124, 512, 1024, 6), (1, 33, 70, 64, 32, 1)])
def test_conv3x3(cuda_ctx, dtype16, n, h, w, cin, cout, dil):
    from marie_icr_b200 import ops
    torch.manual_seed(cin + cout + w)
    x = torch.randn(n, h, w, cin, device="cuda").to(dtype16)
    wt = (torch.randn(cout, cin, 3, 3, device="cuda") * (9 * cin) ** -0.5).to(dtype16)
    bias = torch.randn(cout, device="cuda") * 0.1
    out = ops.conv16(x, ops.pack_conv_weight(wt, dtype16), bias=bias, act=ops.ACT_RELU, taps=9, dil=dil)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=dil, dilation=dil).relu()
    _close(out, ref.permute(0, 2, 3, 1), f"conv3x3 {cin}->{cout} d{dil}")


def test_conv1x1_concat_and_planar(cuda_ctx, dtype16):
    from marie_icr_b200 import ops
    torch.manual_seed(5)
    n, h, w = 1, 31, 150
    x0 = torch.randn(n, h, w, 128, device="cuda").to(dtype16)
    x1 = torch.randn(n, h, w, 64, device="cuda").to(dtype16)
    wt = (torch.randn(64, 192, 1, 1, device="cuda") * 192 ** -0.5).to(dtype16)
    out = ops.conv16(x0, ops.pack_conv_weight(wt, dtype16), x1=x1, taps=1)
    ref = F.conv2d(torch.cat([x0, x1], 3).float().permute(0, 3, 1, 2), wt.float())
    _close(out, ref.permute(0, 2, 3, 1), "conv1x1 concat")
    # small-N head layer written as fp32 planes (rows padded to 16 in the weight matrix)
    w2 = torch.zeros(16, 64, device="cuda")
    w2[:2] = torch.randn(2, 64, device="cuda") * 0.1
    b2 = torch.tensor([0.3, -0.2], device="cuda")
    planes = ops.conv16(x1, w2.to(dtype16), bias=b2, taps=1, n_out=2, planar=True)
    ref2 = F.conv2d(x1.float().permute(0, 3, 1, 2), w2[:2].to(dtype16).float()[:, :, None, None], b2)
    assert planes.shape == (2, 1, h, w)
    assert (planes[:, 0] - ref2[0]).abs().max().item() < 2e-3 * ref2.abs().max().item() + 1e-4


def test_gemm_block_diagonal(cuda_ctx, dtype16):
    """Per-head projections: 16 batches, (K=64 -> N=768) and (K=768 -> N=64) with per-batch column / row offsets."""
    from marie_icr_b200 import ops
    torch.manual_seed(9)
    R, H, E, heads = 300, 1024, 768, 16
    q = torch.randn(R, H, device="cuda").to(dtype16)
    wk = (torch.randn(heads * E, 64, device="cuda") * 0.125).to(dtype16)       # row (h, j): 64 inputs of head h
    qp = ops.gemm16_batched(q, wk, heads, E, 64, 64, E, E, heads * E)
    ref = torch.einsum("rhd,hjd->rhj", q.float().view(R, heads, 64), wk.float().view(heads, E, 64)).reshape(R, heads * E)
    _close(qp, ref, "per-head K-side projection")
    ctx = torch.randn(R, heads * E, device="cuda").to(dtype16)
    wv = (torch.randn(H, E, device="cuda") * E ** -0.5).to(dtype16)
    bias = torch.randn(H, device="cuda")
    att = ops.gemm16_batched(ctx, wv, heads, 64, E, E, 64, 64, H, bias=bias)
    ref2 = torch.einsum("rhe,hde->rhd", ctx.float().view(R, heads, E), wv.float().view(heads, 64, E)).reshape(R, H) + bias
    _close(att, ref2, "per-head V-side projection")


@pytest.mark.parametrize("M,N,K", [(700, 768, 256), (37, 96, 128), (1154, 2304, 768), (300, 80, 64)])
def test_tma_epilogue_equals_register_epilogue(cuda_ctx, dtype16, M, N, K):
    """The TMA-store epilogue (row-per-lane, packed fp32, 64-byte-swizzled staging, ragged edges clipped by the tensor
    map) against the register / fp32-staging epilogue it replaces (MB_EPI_TMA=0): bit-identical for bias / ReLU /
    residual (same operation order), within the output rounding for the re-parametrised GELU; in-place residual too."""
    import os
    from marie_icr_b200 import ops
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device="cuda").to(dtype16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dtype16)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(dtype16)

    def run(**kw):
        outs = []
        for flag in ("1", "0"):
            os.environ["MB_EPI_TMA"] = flag
            try:
                outs.append(ops.gemm16(a, w, **kw))
            finally:
                os.environ.pop("MB_EPI_TMA", None)
        return outs

    for kw in (dict(), dict(bias=bias), dict(bias=bias, act=ops.ACT_RELU), dict(bias=bias, residual=res),
               dict(bias=bias, act=ops.ACT_RELU, residual=res)):
        new, old = run(**kw)
        assert torch.equal(new, old), (sorted(kw), (new.float() - old.float()).abs().max().item())
    new, old = run(bias=bias, act=ops.ACT_GELU)
    ref = F.gelu(a.float() @ w.float().t() + bias)
    _close(new, ref, "gelu (TMA epilogue)")
    assert (new.float() - old.float()).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    # in place: the residual stream is both residual and output (x += proj(...)), as the encoder uses it
    x = res.clone()
    os.environ["MB_EPI_TMA"] = "1"
    try:
        from marie_icr_b200._lib import c_int, c_ll, cur_stream, ptr
        cuda_ctx.call("mb_gemm16", ptr(a), c_ll(K), ptr(w), c_int(N), c_int(M), c_int(N), c_int(K), ptr(bias), c_int(0),
                      ptr(x), c_ll(N), ptr(x), c_ll(N), c_int(0), cur_stream())
    finally:
        os.environ.pop("MB_EPI_TMA", None)
    assert torch.equal(x, ops.gemm16(a, w, bias=bias, residual=res))


@pytest.mark.parametrize("M,N,K", [(577 * 3, 768, 768), (1000, 768, 3072), (130, 1024, 1024), (257, 128, 256), (5, 768, 768)])
def test_residual_gemm_leaves_layernorm_statistics(cuda_ctx, dtype16, M, N, K):
    """The residual GEMM's epilogue sums x and x^2 of the rows it writes (TapGemm::stat_out); the finished (-mean, rstd)
    pairs must describe the stored rows like ln_stats_kernel's two-pass statistics do (encoder proj / fc2 in front of the
    folded LayerNorms)."""
    import torch
    from marie_icr_b200 import ops
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(dtype16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(dtype16)
    bias = torch.randn(N, device="cuda") * 0.1
    res = (torch.randn(M, N, device="cuda") * 2 + 0.7).to(dtype16)        # non-zero row means
    out, stats = ops.gemm16_res_stats(a, w, bias, res)
    ref = a.float() @ w.float().t() + bias + res.float()
    tol = 2e-3 if dtype16 == torch.float16 else 1.6e-2
    assert ((out.float() - ref).norm() / ref.norm()).item() <= tol
    x = out.float()                                                       # statistics of what was stored
    mean, var = x.mean(1), x.var(1, unbiased=False)
    rstd = torch.rsqrt(var + 1e-6)
    assert (stats[:, 0] + mean).abs().max().item() <= 2e-3 * x.abs().max().item()
    assert ((stats[:, 1] - rstd).abs() / rstd).max().item() <= 2e-3
