"""First-contact probe for the tcgen05 tap-GEMM on a real B200: correctness pattern dump + throughput."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marie_icr_b200 import ops  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
res = {}


def check(name, out, ref):
    out = out.float()
    err = (out - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    rel = err.max().item() / scale
    res[name] = rel
    print(f"{name}: rel_err={rel:.3e}", flush=True)
    if rel > 1e-2:
        bad = (err > 1e-2 * scale)
        rows = bad.any(1).nonzero().flatten()[:16].tolist()
        cols = bad.any(0).nonzero().flatten()[:16].tolist()
        print("   bad rows:", rows, "bad cols:", cols, "frac bad:", bad.float().mean().item())
        print("   out[0,:8]", out[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())


torch.manual_seed(0)
for (M, N, K) in [(128, 64, 64), (128, 64, 256), (128, 256, 64), (256, 512, 512), (1731, 768, 768)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    out = ops.gemm_bf16(a, w)
    torch.cuda.synchronize()
    check(f"gemm_{M}x{N}x{K}", out, a.float() @ w.float().t())

import torch.nn.functional as F
x = torch.randn(1, 8, 128, 64, device="cuda").to(torch.bfloat16)
wt = (torch.randn(64, 64, 3, 3, device="cuda") * (9 * 64) ** -0.5).to(torch.bfloat16)
out = ops.conv_bf16(x, ops.pack_conv_weight(wt), taps=9)
torch.cuda.synchronize()
check("conv3x3_64", out.reshape(-1, 64), F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, 64))

# throughput
for (M, N, K) in [(8192, 8192, 8192), (577 * 256, 2304, 768), (577 * 256, 768, 3072)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    for _ in range(3):
        ops.gemm_bf16(a, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm_bf16(a, w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tf = 2.0 * M * N * K / ms / 1e9
    res[f"tflops_{M}x{N}x{K}"] = tf
    print(f"gemm {M}x{N}x{K}: {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)
    t0 = time.time()
    for _ in range(3):
        (a @ w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        (a @ w.t())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[f"cublas_tflops_{M}x{N}x{K}"] = 2.0 * M * N * K / ms / 1e9
    print(f"   cuBLAS: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)

json.dump(res, open("gpurun_out/probe_gemm.json", "w"), indent=1)
