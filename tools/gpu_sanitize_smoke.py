"""Tiny end-to-end pass for compute-sanitizer memcheck: K1 -> CRAFT -> post -> K9 -> TrOCR-tiny encode + greedy + beam."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops, weights
from marie_icr_b200.pipeline import PagePipeline, PSM_PRESETS
from synthetic import pages as synth, weights as sw

page, _ = synth.synth_page(0, height=330, width=255, scale=0.6, line_pitch=40, gap=16, margin=16)
cfg = sw.trocr_tiny()
tsd = sw.synth_trocr_state(cfg, 1)
pipe = PagePipeline(craft_blob=weights.pack_craft(sw.glyph_craft_state(0)), trocr_blob=weights.pack_trocr(tsd, cfg), micro_batch=2,
                    crop_chunk=8)
dev = torch.from_numpy(np.stack([page, page])).cuda()
for beam in (1, 3):
    rec, counts = pipe.run_device(dev, preset=PSM_PRESETS["sparse"], beam=beam, max_len_b=6, out_ld=8)
    torch.cuda.synchronize()
    print("beam", beam, "boxes", counts, "records", tuple(rec.shape))
