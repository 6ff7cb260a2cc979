"""Greedy cross-attention core in isolation: tcgen05 / TMA kernel (mode 0) vs the mma.sync kernel (mode 1).
ROWS crops x 577 encoder tokens x 768; CUDA events, L2 flushed by the working set itself (1.8 GB at 2048 rows).
Prints ms per launch and the algorithmic HBM rate (encoder states + queries + context, each moved once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marie_icr_b200 import ops  # noqa: E402
from marie_icr_b200._lib import Context  # noqa: E402

rows = int(os.environ.get("ROWS", "2048"))
E = int(os.environ.get("EDIM", "768"))
T, heads = 577, 16
Context.get(0)
torch.manual_seed(0)
enc = torch.randn(rows * T, E, device="cuda").half()
qp = (torch.randn(rows, heads * E, device="cuda") * 0.05).half()
for frac in (0.0, 0.5):
    fin = (torch.rand(rows, device="cuda") < frac).to(torch.uint8) if frac else None
    live = rows if fin is None else int((fin == 0).sum())
    for mode in (1, 0, 1, 0):
        for _ in range(3):
            ops.cross_enc16(qp, enc, T, heads, finished=fin, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.cross_enc16(qp, enc, T, heads, finished=fin, mode=mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gb = live * (T * E + 2 * heads * E) * 2 / 1e9
        print(f"finished {frac:.1f} mode {mode} ({'tcgen05' if mode == 0 else 'mma.sync'}): {ms * 1e3:8.1f} us  {gb / ms:7.2f} TB/s"
              f"  ({live} live rows)", flush=True)
