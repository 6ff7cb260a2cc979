"""Recogniser chunk size: throughput and result identity (records must not depend on the chunking)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from marie_icr_b200 import weights
from marie_icr_b200._lib import Context
from marie_icr_b200.pipeline import PSM_PRESETS, PagePipeline

ctx = Context.get(0)
dt = ctx.torch_dtype
npages = int(os.environ.get("NPAGES", 16))
pages_np, _ = bench.make_pages(list(range(npages)))
craft_sd, tsd, cfg = bench.make_weights(dt)
cb, tb = weights.pack_craft(craft_sd, dt), weights.pack_trocr(tsd, cfg, dt)
pages = torch.from_numpy(pages_np).cuda()
kw = dict(preset=PSM_PRESETS["sparse"], beam=1, max_len_b=200, out_ld=32)
ref = None
for chunk in [int(c) for c in os.environ.get("CHUNKS", "2048,4096,8192").split(",")]:
    pipe = PagePipeline(device=0, craft_blob=cb, trocr_blob=tb, micro_batch=8, crop_chunk=chunk)
    rec, counts = pipe.run_device(pages, **kw)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(2): rec, counts = pipe.run_device(pages, **kw)
    torch.cuda.synchronize(); dtm = (time.time() - t0) / 2
    same = None if ref is None else bool(torch.equal(ref, rec))
    if ref is None: ref = rec.clone()
    print(f"chunk {chunk}: {npages / dtm:.2f} pages/s, {int(sum(counts))} crops, identical to first: {same}", flush=True)
    del pipe
