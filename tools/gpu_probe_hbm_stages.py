"""HBM-stage probe (run on the GPU box): K1 page preprocess, K5-K7 post-processing, K9 crop packing on synthetic letter
pages.  Prints stage times (CUDA events, L2 flushed by the size of the inputs) against the algorithmic bytes of
SURVEY.md section 8(d); run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from marie_icr_b200 import ops
from marie_icr_b200._lib import Context
from oracle import synth


def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    Context.get(0)
    npages = int(os.environ.get("NPAGES", 64))
    uniq = min(npages, 8)
    base = [synth.synth_page(i)[0] for i in range(uniq)]
    pages = np.stack([base[i % uniq] for i in range(npages)])
    dpages = torch.from_numpy(pages).cuda()
    mb = 8
    t = timeit(lambda: [ops.page_preprocess(dpages[i:i + mb]) for i in range(0, npages, mb)])
    print(f"K1 preprocess: {t:.3f} ms / {npages} pages = {55.7e-3 * npages / t:.2f} TB/s algorithmic")
    _, ratio = ops.page_preprocess(dpages[:1])
    maps = [synth.score_maps_from_page(base[i], 1280, 992) for i in range(uniq)]
    text = torch.from_numpy(np.stack([maps[i % uniq][0] for i in range(npages)])).cuda()
    link = torch.from_numpy(np.stack([maps[i % uniq][1] for i in range(npages)])).cuda()
    r2 = 2 / ratio
    post = lambda: ops.craft_post(text, link, 0.7, 0.45, 0.3, ratios=[(r2, r2)] * npages, page_hw=[(3300, 2550)] * npages)
    t = timeit(post)
    print(f"K5-K7 post: {t:.3f} ms / {npages} pages = {20.3e-3 * npages / t:.2f} TB/s algorithmic")
    out = post()
    nb = out["n_boxes"].cpu().tolist()
    rects = torch.cat([out["rects"][i, :nb[i]] for i in range(npages)]).contiguous()
    pidx = torch.cat([torch.full((nb[i],), i, dtype=torch.int32, device="cuda") for i in range(npages)])
    n = rects.shape[0]
    chunk = 2048
    src = float(((rects[:, 2] + 1) * (rects[:, 3] + 1) * 3).sum())
    t = timeit(lambda: [ops.pack_crops(dpages, rects[i:i + chunk], pidx[i:i + chunk], layout=1) for i in range(0, n, chunk)], n=3, warm=1)
    print(f"K9 crops: {t:.3f} ms / {n} crops = {(884736.0 * n + src) / t / 1e9:.2f} TB/s algorithmic "
          f"(mean crop {float(rects[:, 2].float().mean()) + 1:.0f} x {float(rects[:, 3].float().mean()) + 1:.0f})")


if __name__ == "__main__":
    main()
