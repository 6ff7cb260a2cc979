"""Sustained (power-capped) throughput of the encoder GEMM shapes: each shape is run back to back for ~0.6 s."""
import os, sys, subprocess, statistics
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
    f(); f(); torch.cuda.synchronize()
    fl = 2.0 * m * n * k
    reps = max(3, int(0.6 / (fl / 0.9e15)))
    mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                           stdout=subprocess.PIPE, text=True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    mon.terminate()
    s = [l.split(",") for l in mon.stdout.read().strip().splitlines() if "," in l]
    clk = statistics.median(float(x) for x, _ in s) if s else 0
    pw = statistics.median(float(y) for _, y in s) if s else 0
    t = e0.elapsed_time(e1) / reps
    print(f"{which} GEMM2={os.environ.get('MB_GEMM2', '0')} {name:9s} N={n:5d} K={k:4d}: {t * 1e3:8.1f} us {fl / t / 1e9:6.0f} TF/s  [{clk:.0f} MHz {pw:.0f} W, {reps} reps]")
    del a, w, res
