"""Per-shape throughput of the tap-GEMM against cuBLAS (torch.matmul) on the shapes the path uses."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops
from marie_icr_b200._lib import Context

ctx = Context.get(0)
dt = ctx.torch_dtype


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


M = 1024 * 577
shapes = [("enc qkv", M, 2304, 768, {}), ("enc proj+res", M, 768, 768, {"res": 1}), ("enc fc1+gelu", M, 3072, 768, {"act": 2, "bias": 1}),
          ("enc fc2+res", M, 768, 3072, {"res": 1, "bias": 1}), ("cross kv", M, 2048, 768, {"bias": 1}),
          ("dec qkv R=1024", 1024, 3072, 1024, {"bias": 1}), ("dec out+res R=1024", 1024, 1024, 1024, {"res": 1, "bias": 1}),
          ("dec fc1 R=1024", 1024, 4096, 1024, {"act": 1, "bias": 1}), ("dec fc2 R=1024", 1024, 1024, 4096, {"res": 1, "bias": 1}),
          ("logits R=1024", 1024, 50265, 1024, {"f32": 1})]
for name, m, n, k, o in shapes:
    a = torch.randn(m, k, device="cuda").to(dt)
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).to(dt)
    bias = torch.randn(n, device="cuda") if o.get("bias") else None
    res = torch.randn(m, n, device="cuda").to(dt) if o.get("res") else None
    f = lambda: ops.gemm16(a, w, bias=bias, act=o.get("act", 0), residual=res, out_dtype=torch.float32 if o.get("f32") else None)
    t = timeit(f)
    tc = timeit(lambda: torch.matmul(a, w.t()))
    fl = 2.0 * m * n * k
    print(f"{name:22s} M={m:7d} N={n:5d} K={k:4d}: ours {t*1e3:8.1f} us {fl/t/1e9:7.0f} TF/s | cuBLAS {tc*1e3:8.1f} us {fl/tc/1e9:7.0f} TF/s")
    del a, w, res
# CRAFT conv layers at 1 page (n=8 pages batch): (cin, cout, h, w, dil)
n_img = 8
convs = [("conv1_2", 64, 64, 2560, 1984, 1), ("conv2_1", 64, 128, 1280, 992, 1), ("conv2_2", 128, 128, 1280, 992, 1),
         ("conv3_2", 256, 256, 640, 496, 1), ("conv4_2", 512, 512, 320, 248, 1), ("conv5_1", 512, 512, 160, 124, 1),
         ("fc6 d6", 512, 1024, 160, 124, 6), ("cls1 (64->64 pad)", 64, 64, 1280, 992, 1)]
for name, ci, co, h, w_, dil in convs:
    x = torch.randn(n_img, h, w_, ci, device="cuda").to(dt)
    wt = (torch.randn(co, 9 * ci, device="cuda") * (9 * ci) ** -0.5).to(dt)
    b = torch.zeros(co, device="cuda")
    t = timeit(lambda: ops.conv16(x, wt, bias=b, act=1, taps=9, dil=dil), n=3)
    fl = 2.0 * n_img * h * w_ * co * 9 * ci
    print(f"{name:22s} {ci:4d}->{co:4d} @{h}x{w_} x{n_img}: {t*1e3:8.1f} us {fl/t/1e9:7.0f} TF/s")
    del x, wt
