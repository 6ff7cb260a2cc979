"""Small run of the memory-bound stages (K1, K5-K7, K9) for ncu launch lists."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marie_icr_b200 import ops
from marie_icr_b200._lib import Context
from synthetic import pages as synth

n = int(os.environ.get("NPAGES", 8))
ctx = Context.get(0)
pages = np.stack([synth.synth_page(i)[0] for i in range(n)])
dpages = torch.from_numpy(pages).cuda()
maps = [synth.score_maps_from_page(pages[i], 1280, 992) for i in range(n)]
text = torch.from_numpy(np.stack([m[0] for m in maps])).cuda()
link = torch.from_numpy(np.stack([m[1] for m in maps])).cuda()
r2 = 2 / (2550 / 3300)
for it in range(2):
    x, ratio = ops.page_preprocess(dpages)
    out = ops.craft_post(text, link, 0.7, 0.45, 0.3, ratios=[(r2, r2)] * n, page_hw=[(3300, 2550)] * n)
    nb = out["n_boxes"].cpu().tolist()
    rects = torch.cat([out["rects"][i, :nb[i]] for i in range(n)]).contiguous()
    pidx = torch.cat([torch.full((nb[i],), i, dtype=torch.int32, device="cuda") for i in range(n)])
    patches = ops.pack_crops(dpages, rects, pidx, layout=1)
torch.cuda.synchronize()
print("ok", sum(nb), ctx.launches)
