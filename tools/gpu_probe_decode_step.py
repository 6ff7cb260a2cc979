"""Greedy decode of NCROPS random encoder states for a few steps (ncu launch lists of the decode loop at batch size)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops, weights
from marie_icr_b200._lib import Context
from oracle import trocr

n = int(os.environ.get("NCROPS", 8192))
ctx = Context.get(0)
dt = ctx.torch_dtype
cfg = trocr.trocr_base()
sd = trocr.synth_trocr_state(cfg, 0, round_to=dt)          # no EOS calibration: every crop stays live
ops.load_trocr(weights.pack_trocr(sd, cfg, dt))
torch.manual_seed(0)
enc = (torch.randn(n, 577, 768, device="cuda") * 0.5).to(dt)
_, lens, _, steps = ops.trocr_decode(enc, beam=1, max_len_b=int(os.environ.get("MAXLEN", 3)))
torch.cuda.synchronize()
print("ok", steps, ctx.launches)
