"""Encoder attention micro-benchmark (run on the GPU box): attn_tc_kernel over N crops x 12 heads x 577 tokens."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops
from marie_icr_b200._lib import Context


def main():
    ctx = Context.get(0)
    dt = ctx.torch_dtype
    n, T, heads = int(os.environ.get("NCROPS", 2048)), 577, 12
    qkv = (torch.randn(n * T, 3 * heads * 64, device="cuda") * 1.5).to(dt)
    for mode in (0,):
        for _ in range(2): ops.attention16(qkv, n, T, 0.125, mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        reps = 5
        e0.record()
        for _ in range(reps): ops.attention16(qkv, n, T, 0.125, mode)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 4.0 * T * T * 64 * heads * n
        exps = n * heads * 640 * 640
        print(f"attention mode {mode}: {ms:.3f} ms for {n} crops = {flops / ms / 1e9:.0f} TFLOP/s, "
              f"{exps / ms / 1e6 / 148:.2f} ex2/ns/SM (MUFU peak 16/clk)")


if __name__ == "__main__":
    main()
