"""Burst (not power-capped) timing of the encoder GEMM shapes: one launch after an idle gap, median of 7."""
import os, sys, time, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops
from marie_icr_b200._lib import Context

ctx = Context.get(0)
dt = ctx.torch_dtype
M = int(os.environ.get("NCROPS", 2048)) * 577
shapes = [("qkv", M, 2304, 768, {}), ("proj+res", M, 768, 768, {"res": 1, "bias": 1}), ("fc1+gelu", M, 3072, 768, {"act": 2, "bias": 1}),
          ("fc2+res", M, 768, 3072, {"res": 1, "bias": 1})]
which = os.environ.get("WHICH", "ours")
for name, m, n, k, o in shapes:
    a = (torch.randn(m, k, device="cuda") * 0.5).to(dt)
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).to(dt)
    bias = torch.randn(n, device="cuda") if o.get("bias") else None
    res = torch.randn(m, n, device="cuda").to(dt) if o.get("res") else None
    if which == "cublas":
        out = torch.empty(m, n, device="cuda", dtype=dt)
        f = lambda: torch.matmul(a, w.t(), out=out)
    else:
        f = lambda: ops.gemm16(a, w, bias=bias, act=o.get("act", 0), residual=res)
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        time.sleep(0.4)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = statistics.median(ts)
    fl = 2.0 * m * n * k
    print(f"burst {which} GEMM2={os.environ.get('MB_GEMM2', '0')} {name:9s}: {t * 1e3:8.1f} us {fl / t / 1e9:6.0f} TF/s")
    del a, w, res
