"""Greedy decode step timing at chunk size (run on the GPU box): ms per decoder step for N crops of random encoder states,
EOS suppressed (every crop stays live), and the implied encoder-state bandwidth of the cache-free cross-attention."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from marie_icr_b200 import ops, weights
from marie_icr_b200._lib import Context
from oracle import trocr

ctx = Context.get(0)
dt = ctx.torch_dtype
cfg = trocr.trocr_base()
sd = trocr.synth_trocr_state(cfg, 0, round_to=dt)          # no EOS calibration: hypotheses run to max_len
ops.load_trocr(weights.pack_trocr(sd, cfg, dt))
for n in (2048,):
    enc = (torch.randn(n, 577, 768, device="cuda") * 0.5).to(dt)
    ops.trocr_decode(enc, beam=1, max_len_b=4)
    res = {}
    for ml in (3, 11):
        best = 1e9
        for _ in range(4):                                   # best of 4: the wall-clock of one short decode is noisy
            torch.cuda.synchronize(); t0 = time.time()
            _, lens, _, steps = ops.trocr_decode(enc, beam=1, max_len_b=ml)
            torch.cuda.synchronize(); best = min(best, (time.time() - t0) * 1e3)
        res[ml] = (best, steps)
    per = (res[11][0] - res[3][0]) / (res[11][1] - res[3][1])
    gb = 12 * n * 577 * 768 * 2 / 1e9
    print(f"n={n}: {per:.2f} ms/step ({res[11][1]} steps, mean len {float(lens.float().mean()):.1f}); encoder states read per step "
          f"{gb:.1f} GB -> if cross-attention were the whole step: {gb / per:.2f} TB/s")
