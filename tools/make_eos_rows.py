"""Pre-computes the calibrated EOS row of the synthetic TrOCR output projection (synthetic/eos_row_<name>.npy) with the
oracle's calibrate_eos on real encoder states of synthetic word crops, so that bench.py's GPU arm never calls into
oracle/.  Run in the build container:   python tools/make_eos_rows.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resample, trocr  # noqa: E402
from synthetic import pages, weights  # noqa: E402


def calib_fragments(page, k=8):
    return [page[145 + 70 * j:205 + 70 * j, 150 + 90 * (j % 3):370 + 140 * (j % 4)].copy() for j in range(k)]


def test_fragments(page, k=12):
    """Other crops (different words / sizes) to check that every hypothesis terminates early."""
    rng = np.random.default_rng(0)
    out = []
    for _ in range(k):
        y, x = int(rng.integers(150, 2900)), int(rng.integers(150, 2000))
        out.append(page[y:y + int(rng.integers(40, 80)), x:x + int(rng.integers(80, 320))].copy())
    return out


def main():
    page, _ = pages.synth_page(0)
    frags = calib_fragments(page)
    cal = torch.stack([torch.from_numpy(resample.fragment_to_input(f)) for f in frags])
    for name, cfg, seed, step in (("trocr_base_seed0", weights.trocr_base(), 0, 5), ("trocr_large_seed0", weights.trocr_large(), 0, 5)):
        sd = weights.synth_trocr_state(cfg, seed, round_to=None)
        with torch.no_grad():
            enc = trocr.encoder_forward(sd, cfg, cal)
            alpha = trocr.calibrate_eos(sd, cfg, eos_step=step, round_to=None, enc=enc, margin=1.0)
            more = torch.stack([torch.from_numpy(resample.fragment_to_input(f)) for f in test_fragments(page)])
            hyps = trocr.generate(sd, cfg, trocr.encoder_forward(sd, cfg, more), beam=1, max_len_b=200)
        row = sd["decoder.output_projection.weight"][weights.EOS].numpy().astype(np.float32)
        np.save(os.path.join(ROOT, "synthetic", f"eos_row_{name}.npy"), row)
        print(name, "alpha", alpha, "greedy lengths on 12 other crops:", [len(h[0]["tokens"]) for h in hyps])


if __name__ == "__main__":
    main()
