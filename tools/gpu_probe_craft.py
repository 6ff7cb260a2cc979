"""CRAFT forward probe (run on the GPU box, plain or under ncu): NPAGES letter pages through K1 + mb_craft_forward.
Plain run prints ms/page and TFLOP/s; under `ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...`
with `-k regex:'tap_gemm|conv1_1|maxpool|upsample'` the launch list is the per-layer table (tools/ncu_craft_table.py)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops, weights
from marie_icr_b200._lib import Context
from synthetic import pages as synth, weights as sw

ctx = Context.get(0)
dt = ctx.torch_dtype
npages = int(os.environ.get("NPAGES", 8))
reps = int(os.environ.get("REPS", 3))
pages = torch.from_numpy(np.stack([synth.synth_page(i)[0] for i in range(npages)])).cuda()
ops.load_craft(weights.pack_craft(sw.glyph_craft_state(0), dt))
x, ratio = ops.page_preprocess(pages)
ops.craft_forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps):
    ops.craft_forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps / npages
print(f"CRAFT forward, {npages} pages/launch set: {ms:.3f} ms/page = {3.613 / ms * 1e3:.0f} TFLOP/s algorithmic")
