"""TrOCR-base encoder timing (run on the GPU box): N crops of random patches through mb_trocr_encode."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from marie_icr_b200 import ops, weights
from marie_icr_b200._lib import Context
from oracle import trocr


def main():
    ctx = Context.get(0)
    dt = ctx.torch_dtype
    cfg = trocr.trocr_base()
    sd = trocr.synth_trocr_state(cfg, 0, round_to=dt)
    ops.load_trocr(weights.pack_trocr(sd, cfg, dt))
    n = int(os.environ.get("NCROPS", 2048))
    patches = (torch.randn(n * 576, 768, device="cuda") * 0.5).to(dt)
    for _ in range(2): ops.trocr_encode(patches)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = int(os.environ.get("REPS", 6))
    import subprocess, statistics
    mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                           stdout=subprocess.PIPE, text=True)
    e0.record()
    for _ in range(reps): out = ops.trocr_encode(patches)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    mon.terminate()
    samples = [l.split(",") for l in mon.stdout.read().strip().splitlines() if "," in l]
    q = "no samples"
    if samples:
        q = f"{statistics.median(float(a) for a, _ in samples):.0f} MHz, {statistics.median(float(b) for _, b in samples):.0f} W median of {len(samples)}"
    print(f"encoder n={n} LNFOLD={os.environ.get('MB_LNFOLD', '1')} skip={os.environ.get('MB_PROBE_SKIP', '-')} [{q}]: {ms:.2f} ms = {ms / n * 1e3:.1f} us/crop "
          f"({111e9 * n / (ms * 1e-3) / 1e12:.0f} TFLOP/s), finite={bool(torch.isfinite(out.float()).all())}")


if __name__ == "__main__":
    main()
