"""bench.py — the hot path on synthetic letter pages: CRAFT box detection + TrOCR-base greedy ICR (BASELINE.json
configs[1]: batch of 64 pages per GPU).  One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (B200, hand-written CUDA through the C ABI)
  python bench.py --impl reference [...]                         the reference's algorithm on the host cores
                                                                 (oracle port: the reference needs fairseq/timm, absent)
A "step" is one pass of the whole path over one batch of pages per GPU.  `value` = pages/s with the page batch already
resident in HBM; `e2e` = the same through host buffers (pinned pages -> device, word records -> host) every step.
Under torchrun each rank runs its own batch (weak scaling: page i of the stream -> rank i mod world) and the packed
word records are gathered with one NCCL all_gather pair inside the timed region.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "pages/sec (CRAFT detect + TrOCR-base greedy ICR, 2550x3300 synthetic letter pages, ~510 word crops/page)"
OUT_LD = 32           # tokens kept per word record
MAX_LEN_B = 200       # task.py:266


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pages", type=int, default=64, help="pages per GPU per step (configs[1]: 64)")
    ap.add_argument("--beam", type=int, default=1)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--crop-chunk", type=int, default=16384, help="max crops per decode batch (results do not depend on it)")
    ap.add_argument("--encode-chunk", type=int, default=2048, help="crops per K9 + encoder pass inside a decode batch")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------- inputs
def make_pages(indices):
    from synthetic import pages as synth
    pages, words = [], 0
    for i in indices:
        p, w = synth.synth_page(i)
        pages.append(p)
        words += w
    return np.stack(pages), words


def make_weights(dtype):
    """Seeded synthetic weights shared by both arms (synthetic/weights.py — data generators, not oracle code):
    glyph-path CRAFT (text-like maps) and TrOCR-base with the pre-computed calibrated EOS row (hypotheses end after
    ~6 tokens).  dtype: the 16-bit type the weights are rounded to once (None = keep fp32).  -> (craft_sd, trocr_sd, cfg)"""
    from synthetic import weights as sw
    craft_sd = sw.glyph_craft_state(0)
    cfg = sw.trocr_base()
    tsd = sw.apply_eos_row(sw.synth_trocr_state(cfg, 0, round_to=dtype), "trocr_base_seed0", round_to=dtype)
    return craft_sd, tsd, cfg


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(page, craft_sd, tsd, cfg, beam, strip_rows=1650, n_crops=24):
    """(The only place bench.py executes oracle/ code.)  One bounded sample of the reference's CPU path on one page: K1 + CRAFT.forward + getDetBoxes + crop/resample on a
    horizontal strip of `strip_rows` page rows, TrOCR (encoder + search) on `n_crops` of the strip's crops, both
    extrapolated to the full page.  Returns (seconds per page, crops per page estimate, detail dict)."""
    from oracle import craft_net, craft_post, resample, trocr
    ph = page.shape[0]
    strip = np.ascontiguousarray(page[150:150 + strip_rows])
    t0 = time.perf_counter()
    x, ratio = resample.craft_input(strip)
    xin = torch.from_numpy(x).permute(2, 0, 1)[None]
    with torch.no_grad():
        y, _ = craft_net.craft_forward(craft_sd, xin)
    t1 = time.perf_counter()
    det, _, _ = craft_post.det_boxes_cv(y[0, ..., 0].numpy(), y[0, ..., 1].numpy(), 0.7, 0.45, 0.3)
    adj = craft_post.adjust_result_coordinates([b.copy() for b in det], 1 / ratio, 1 / ratio)
    rects = craft_post.boxes_to_rects(adj, strip.shape[0], strip.shape[1])
    frags = [craft_post.crop_rect(strip, r) for r in rects]
    t2 = time.perf_counter()
    sample = frags[:n_crops] if frags else []
    if sample:
        chw = torch.stack([torch.from_numpy(resample.fragment_to_input(f)) for f in sample])
        with torch.no_grad():
            trocr.recognize(tsd, cfg, chw, beam=beam, max_len_b=MAX_LEN_B)
    t3 = time.perf_counter()
    scale = ph / strip_rows
    crops_page = len(frags) * scale
    per_crop = (t3 - t2) / max(len(sample), 1)
    sec_page = (t1 - t0) * scale + (t2 - t1) * scale + per_crop * crops_page
    return sec_page, crops_page, dict(craft_s=(t1 - t0) * scale, post_crop_s=(t2 - t1) * scale, trocr_s_per_crop=per_crop)


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    pages, _ = make_pages([0])
    craft_sd, tsd, cfg = make_weights(None)
    secs, crops = [], []
    for i in range(args.warmup + args.steps):
        s, c, detail = cpu_reference_step(pages[0], craft_sd, tsd, cfg, args.beam)
        if i >= args.warmup:
            secs.append(s)
            crops.append(c)
    sec_page = float(np.mean(secs))
    value = 1.0 / sec_page
    sample = ("per step: K1+CRAFT.forward+getDetBoxes+crops on a 1650-row strip (half) of one letter page, TrOCR-base (fp32) on 24 "
              "of its crops; extrapolated x2 rows and to all crops of the page")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_page * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 64 synthetic letter pages/GPU/step, CRAFT detect + TrOCR-base greedy ICR",
                   "pages_per_step": 1, "beam": args.beam, "crops_per_page": float(np.mean(crops))},
        "cpu_baseline": {"value": value, "unit": "pages/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                         "detail": detail},
        "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "crops_per_s": value * float(np.mean(crops)),
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from marie_icr_b200 import weights
    from marie_icr_b200._lib import Context
    from marie_icr_b200.dist import gather_records, shard_indices
    from marie_icr_b200.pipeline import PSM_PRESETS, RECORD_HEAD, PagePipeline

    torch.cuda.set_device(local_rank)
    ctx = Context.get(local_rank)
    ctx.set_dtype(args.dtype)
    dt = ctx.torch_dtype
    # page i of the stream -> rank i mod world; every rank holds `pages` pages per step (weak scaling)
    idx = shard_indices(args.pages * world, rank, world)
    pages_np, _ = make_pages(idx)
    craft_sd, tsd, cfg = make_weights(dt)
    pipe = PagePipeline(device=local_rank, craft_blob=weights.pack_craft(craft_sd, dt),
                        trocr_blob=weights.pack_trocr(tsd, cfg, dt), micro_batch=8, crop_chunk=args.crop_chunk, encode_chunk=args.encode_chunk)
    pages_host = torch.from_numpy(pages_np).pin_memory()
    pages_dev = pages_host.cuda(non_blocking=True)
    page_ids = torch.tensor(idx, dtype=torch.int32, device="cuda")
    kw = dict(preset=PSM_PRESETS["sparse"], beam=args.beam, max_len_b=MAX_LEN_B, out_ld=OUT_LD)

    def step_device():
        rec, counts = pipe.run_device(pages_dev, **kw)
        if rec.shape[0]:
            rec[:, 0] = page_ids[rec[:, 0].long()]          # local page slot -> global page id
        return gather_records(rec) if world > 1 else rec, counts

    def step_host():
        dev = pages_host.cuda(non_blocking=True)
        rec, counts = pipe.run_device(dev, **kw)
        if rec.shape[0]:
            rec[:, 0] = page_ids[rec[:, 0].long()]
        rec = gather_records(rec) if world > 1 else rec
        return rec.cpu(), counts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    from marie_icr_b200 import ops as _ops
    st0 = _ops.trocr_stats(local_rank)
    ctx.profile(True)
    pipe.timer.reset(True)
    ms, (rec, counts) = timed(step_device, args.steps)
    stages = pipe.timer.collect()
    pipe.timer.reset(False)
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    st1 = _ops.trocr_stats(local_rank)
    dec_steps = (st1["decode_steps"] - st0["decode_steps"]) / max(1, st1["decode_calls"] - st0["decode_calls"])
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, (rec_h, _) = timed(step_host, max(1, min(args.steps, 2)))
    e2e_steps = max(1, min(args.steps, 2))

    n_crops_local = int(sum(counts))
    tot = torch.tensor([n_crops_local], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tot)
    crops_step = int(tot.item())
    pages_step = args.pages * world
    sec_step = ms / 1e3 / args.steps
    value = pages_step / sec_step
    sec_e2e = ms_e2e / 1e3 / e2e_steps
    if rank != 0:
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json (sustained)" if peaks else "fallback"
    achieved = prof["flops"] / (prof["ms"] * 1e-3) / 1e12 if prof["ms"] > 0 else 0.0
    # DRAM traffic of the dominant tap-GEMM launch (encoder fc1) from the committed ncu --set full capture
    traffic, traffic_detail = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_tap_gemm_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_per_launch"]
        traffic_detail = {"unit": "bytes per launch (dram read + write)", "launch": tj["kernel"],
                          "algorithmic_bytes_per_launch": tj["algorithmic_bytes_per_launch"],
                          "source": "profiles/r01_tap_gemm_traffic.json"}
    except Exception:
        pass
    st_ms, st_units = stages["ms"], stages["units"]
    per_step = {k: v / args.steps for k, v in st_ms.items()}
    # algorithmic bytes per unit: K1 55.7 MB/page, K5-K7 20.3 MB/page, K9 ~0.9 MB/crop
    hbm = {}
    for name, bytes_unit in (("k1_preprocess", 55.72e6), ("k5_7_post", 20.32e6), ("k9_crops", 0.90e6)):
        if st_ms.get(name):
            gbs = bytes_unit * st_units[name] / (st_ms[name] * 1e-3) / 1e9
            hbm[name] = {"achieved_gbs": gbs, "frac": gbs / hbm_peak}
    line = {
        "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "configs[1]: 64 synthetic letter pages/GPU/step, CRAFT detect + TrOCR-base greedy ICR",
                   "pages_per_step": pages_step, "page": "2550x3300x3 u8", "crops_per_step": crops_step, "beam": args.beam,
                   "max_len_b": MAX_LEN_B, "trocr": "base (ViT 768/12 + decoder 1024/12, vocab 50265)",
                   "weights": "seeded synthetic (glyph-path CRAFT, EOS-calibrated TrOCR)", "parallelism": f"dp{world}",
                   "l2": "inputs larger than L2 (1.6 GB of pages per step)", "decoder_steps_per_chunk": dec_steps},
        "crops_per_s": crops_step / sec_step,
        "e2e": {"value": pages_step / sec_e2e, "unit": "pages/s", "h2d_bytes_per_step": int(pages_host.numel()),
                "d2h_bytes_per_step": int(rec_h.numel() * 4 / max(world, 1)), "crops_per_s": crops_step / sec_e2e},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "tap_gemm_kernel (tcgen05 implicit-GEMM conv / linear)", "bound": "tensor",
                     "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak, "traffic": traffic,
                     "traffic_detail": traffic_detail, "peak_source": peak_src, "launches": prof["launches"],
                     "share_of_step": prof["ms"] / ms if ms else None},
        "stages_ms_per_step": per_step, "hbm_stages": hbm,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        craft32, tsd32, cfg32 = make_weights(None)
        t0 = time.perf_counter()
        s, c, detail = cpu_reference_step(pages_np[0], craft32, tsd32, cfg32, args.beam)
        line["cpu_baseline"] = {"value": 1.0 / s, "unit": "pages/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "one pass: K1+CRAFT+getDetBoxes+crops on a 1650-row strip (half) of one page, TrOCR-base "
                                          "fp32 on 24 crops, extrapolated to the page (%.1f s measured)" % (time.perf_counter() - t0),
                                "detail": detail}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 during the run (NCCL prints its version
    banner there) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
