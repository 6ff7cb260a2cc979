"""Stage timing probe (run on the GPU box): CRAFT forward / post / crops / TrOCR encoder / decoder steps."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from marie_icr_b200 import ops, weights
from marie_icr_b200._lib import Context
from oracle import craft_net, synth, trocr


def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ctx = Context.get(0)
    dt = ctx.torch_dtype
    npages = int(os.environ.get("NPAGES", 4))
    pages = np.stack([synth.synth_page(i)[0] for i in range(npages)])
    dpages = torch.from_numpy(pages).cuda()
    ops.load_craft(weights.pack_craft(craft_net.synth_craft_state(0), dt))
    t = timeit(lambda: ops.page_preprocess(dpages))
    print(f"K1 preprocess: {t / npages:.3f} ms/page  ({55.7e-3 / (t / npages) :.0f} GB/s algorithmic)")
    x, ratio = ops.page_preprocess(dpages)
    t = timeit(lambda: ops.craft_forward(x), n=2)
    print(f"CRAFT forward: {t / npages:.2f} ms/page  ({3.613 / (t / npages) * 1e3:.0f} TFLOP/s)")
    maps = [synth.score_maps_from_page(pages[i], 1280, 992) for i in range(npages)]
    text = torch.from_numpy(np.stack([m[0] for m in maps])).cuda()
    link = torch.from_numpy(np.stack([m[1] for m in maps])).cuda()
    r2 = 2 / ratio
    post = lambda: ops.craft_post(text, link, 0.7, 0.45, 0.3, ratios=[(r2, r2)] * npages, page_hw=[(3300, 2550)] * npages)
    t = timeit(post)
    print(f"K5-K7 post: {t / npages:.3f} ms/page ({20.3e-3 / (t / npages):.0f} GB/s algorithmic)")
    out = post()
    nb = out["n_boxes"].cpu().tolist()
    print("boxes per page", nb)
    rects = torch.cat([out["rects"][i, :nb[i]] for i in range(npages)]).contiguous()
    pidx = torch.cat([torch.full((nb[i],), i, dtype=torch.int32, device="cuda") for i in range(npages)])
    n = rects.shape[0]
    t = timeit(lambda: ops.pack_crops(dpages, rects, pidx, layout=1))
    print(f"K9 crops: {t / n * 1e3:.2f} us/crop ({0.9e-3 / (t / n):.0f} GB/s algorithmic), n={n}")
    patches = ops.pack_crops(dpages, rects, pidx, layout=1)
    xln = torch.randn(1024 * 577, 768, device="cuda").to(dt)
    gl, bl = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
    t = timeit(lambda: ops.layernorm16(xln, gl, bl), n=5)
    print(f"LayerNorm 590848x768: {t * 1e3:.0f} us ({xln.numel() * 4 / t / 1e9:.2f} TB/s)")
    del xln
    cfg = trocr.trocr_base()
    t0 = time.time()
    sd = trocr.synth_trocr_state(cfg, 0, round_to=dt)
    blob = weights.pack_trocr(sd, cfg, dt)
    print(f"weights {time.time() - t0:.1f}s, blob {len(blob) / 1e6:.0f} MB")
    ops.load_trocr(blob)
    for nn in (256, 1024):
        nn = min(nn, n)
        p = patches[: nn * 576]
        t = timeit(lambda: ops.trocr_encode(p), n=2)
        print(f"encoder n={nn}: {t / nn * 1e3:.1f} us/crop ({111e9 * nn / (t * 1e-3) / 1e12:.0f} TFLOP/s)")
    enc = ops.trocr_encode(patches[: min(n, 512) * 576])
    for beam in (1, 5):
        ops.trocr_decode(enc, beam=beam, max_len_b=8)          # warm-up: workspace allocation
        res = {}
        for ml in (2, 8):
            torch.cuda.synchronize(); t0 = time.time()
            toks, lens, sc, steps = ops.trocr_decode(enc, beam=beam, max_len_b=ml)
            torch.cuda.synchronize(); res[ml] = ((time.time() - t0) * 1e3, steps)
        per_step = (res[8][0] - res[2][0]) / (res[8][1] - res[2][1])
        print(f"decode beam={beam} n={enc.shape[0]}: {res[8][0]:.1f} ms for {res[8][1]} steps; {per_step:.2f} ms/step, "
              f"prepare (cross K/V) {res[2][0] - per_step * res[2][1]:.1f} ms")
    print("launches", ctx.launches)


if __name__ == "__main__":
    main()
