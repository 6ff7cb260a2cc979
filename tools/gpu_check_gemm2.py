"""Correctness + speed of the two-CTA (cta_group::2) tap-GEMM path (MB_GEMM2=1) on encoder-sized problems."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops
from marie_icr_b200._lib import Context
ctx = Context.get(0)
dt = ctx.torch_dtype
torch.manual_seed(0)
for (M, N, K, act, res) in [(40000, 768, 768, 0, True), (40001, 2304, 768, 0, False), (38000, 3072, 768, 2, False), (45000, 768, 3072, 0, True),
                            (37900, 1000, 256, 1, False)]:
    a = torch.randn(M, K, device="cuda").to(dt)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").to(dt) if res else None
    out = ops.gemm16(a, w, bias=bias, act=act, residual=r).float()
    ref = a.float() @ w.float().t() + bias
    if act == 1: ref = ref.relu()
    if act == 2: ref = F.gelu(ref)
    if res: ref = ref + r.float()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"M={M} N={N} K={K} act={act} res={res}: max rel err {err:.2e}", "OK" if err < 1e-2 else "FAIL", flush=True)
torch.cuda.synchronize()
print("done")
