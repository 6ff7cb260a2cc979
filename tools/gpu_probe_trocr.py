"""Small TrOCR-base run for ncu launch lists: encode N crops, decode a few steps."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from marie_icr_b200 import ops, weights
from marie_icr_b200._lib import Context
from oracle import trocr

n = int(os.environ.get("NCROPS", 128))
beam = int(os.environ.get("BEAM", 1))
ctx = Context.get(0)
dt = ctx.torch_dtype
cfg = trocr.trocr_base()
sd = trocr.synth_trocr_state(cfg, 0, round_to=dt)
ops.load_trocr(weights.pack_trocr(sd, cfg, dt))
torch.manual_seed(0)
patches = (torch.rand(n * 576, 768, device="cuda") * 2 - 1).to(dt)
enc = ops.trocr_encode(patches)
toks, lens, sc, steps = ops.trocr_decode(enc, beam=beam, max_len_b=3)
torch.cuda.synchronize()
print("ok", steps, ctx.launches)
