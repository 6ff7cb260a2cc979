"""Debug aid: which stage of detection depends on a page's position in the batch?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from marie_icr_b200 import ops, weights
from marie_icr_b200._lib import Context
from synthetic import pages as synth, weights as sw

ctx = Context.get(0)
dt = ctx.torch_dtype
ops.load_craft(weights.pack_craft(sw.glyph_craft_state(0), dt))
pg = [synth.synth_page(i)[0] for i in range(3)]
order = [0, 1, 0, 2]
perm = [3, 2, 1, 0]
def run(idx):
    batch = torch.from_numpy(np.stack([pg[i] for i in idx])).cuda()
    x, ratio = ops.page_preprocess(batch)
    sc = ops.craft_forward(x)
    r2 = 2 / ratio
    out = ops.craft_post(sc[0].contiguous(), sc[1].contiguous(), 0.7, 0.45, 0.3, ratios=[(r2, r2)] * 4, page_hw=[(3300, 2550)] * 4)
    torch.cuda.synchronize()
    return x, sc, out
for trial in range(8):
    xa, sa, oa = run(order)
    xb, sb, ob = run([order[p] for p in perm])
    for new, old in enumerate(perm):
        k1 = torch.equal(xa[old], xb[new])
        cr = torch.equal(sa[:, old], sb[:, new])
        lab = torch.equal(oa["labels"][old], ob["labels"][new])
        nl = (int(oa["n_labels"][old]), int(ob["n_labels"][new]))
        nb = (int(oa["n_boxes"][old]), int(ob["n_boxes"][new]))
        st = torch.equal(oa["stats"][old, :nl[0]], ob["stats"][new, :nl[1]]) if nl[0] == nl[1] else False
        print(f"trial {trial} page slot {old}->{new}: K1 {k1} CRAFT {cr} labels {lab} n_labels {nl} stats {st} n_boxes {nb}")
        if nl[0] == nl[1] and not st:
            a, b = oa["stats"][old, :nl[0]].cpu(), ob["stats"][new, :nl[1]].cpu()
            bad = (a != b).any(1).nonzero().flatten()[:5]
            for i in bad.tolist(): print("   label", i, a[i].tolist(), b[i].tolist())
