"""Greedy cross-attention core (per-head projected queries over the encoder states): the tcgen05 / TMA kernel
(xattn_tc.cu, mode 0) and the mma.sync kernel it replaces (mode 1) against an fp32 torch reference of the same op
(softmax_t(q'^h . e_t) . e per crop and head — fairseq's decoder cross-attention with the K / V projections hoisted
out, marie/models/unilm/trocr/trocr_models.py:142-147).  Tolerance: relative L2 <= 3e-3 (fp16) / 2e-2 (bf16): P is
rounded to the 16-bit operand type in both kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qp, enc, rows, T, heads, beam=1):
    E = enc.shape[1]
    q = qp.float().view(rows, beam * heads, E)                 # the hypotheses of a crop attend over the same states
    e = enc.float().view(rows, T, E)
    p = torch.softmax(torch.einsum("rhe,rte->rht", q, e), -1)
    return torch.einsum("rht,rte->rhe", p, e).reshape(rows * beam, heads * E)


def _inputs(rows, T, heads, E, dtype, seed, qscale=0.05, beam=1):
    torch.manual_seed(seed)
    enc = torch.randn(rows * T, E, device="cuda").to(dtype)
    qp = (torch.randn(rows * beam, heads * E, device="cuda") * qscale).to(dtype)
    return qp, enc


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("rows,T,heads,E", [(5, 577, 16, 768), (3, 577, 16, 1024), (9, 50, 2, 128), (2, 64, 16, 768),
                                            (2, 65, 16, 768), (1, 1, 16, 768), (301, 130, 16, 768), (4, 577, 12, 768)])
def test_cross_enc_matches_reference(cuda_ctx, dtype16, mode, rows, T, heads, E):
    from marie_icr_b200 import ops
    qp, enc = _inputs(rows, T, heads, E, dtype16, rows * T + heads)
    out = ops.cross_enc16(qp, enc, T, heads, mode=mode).float()
    ref = _ref(qp, enc, rows, T, heads)
    assert torch.isfinite(out).all()
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel <= (3e-3 if dtype16 == torch.float16 else 2e-2), rel
    worst = ((out - ref).view(rows, -1).norm(dim=1) / ref.view(rows, -1).norm(dim=1)).max().item()
    assert worst <= (6e-3 if dtype16 == torch.float16 else 4e-2), worst      # no single crop is off


@pytest.mark.parametrize("mode", [0, 1])
def test_cross_enc_growing_scores_force_the_rescale_path(cuda_ctx, dtype16, mode):
    """Scores that grow along the sequence by far more than 2^8 per key tile: the tcgen05 kernel has to move its per-head
    references (and rescale the TMEM accumulators) in nearly every tile; rows dominated by one key."""
    from marie_icr_b200 import ops
    rows, T, heads, E = 6, 577, 16, 768
    torch.manual_seed(3)
    enc = torch.randn(rows * T, E, device="cuda")
    ramp = torch.linspace(0.2, 3.0, T, device="cuda").repeat(rows)[:, None]
    enc = (enc.abs() * ramp).to(dtype16)
    qp = (torch.rand(rows, heads * E, device="cuda") * 0.06).to(dtype16)     # positive queries: later keys score much higher
    out = ops.cross_enc16(qp, enc, T, heads, mode=mode).float()
    ref = _ref(qp, enc, rows, T, heads)
    assert torch.isfinite(out).all()
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel <= (4e-3 if dtype16 == torch.float16 else 2e-2), rel


@pytest.mark.parametrize("mode", [0, 1])
def test_cross_enc_skips_finished_rows(cuda_ctx, mode):
    from marie_icr_b200 import ops
    rows, T, heads, E = 333, 70, 16, 768
    qp, enc = _inputs(rows, T, heads, E, torch.float16, 11)
    torch.manual_seed(5)
    fin = (torch.rand(rows, device="cuda") < 0.4).to(torch.uint8)
    out = ops.cross_enc16(qp, enc, T, heads, finished=fin, mode=mode).float()
    ref = _ref(qp, enc, rows, T, heads)
    live = fin == 0
    rel = ((out[live] - ref[live]).norm() / ref[live].norm()).item()
    assert rel <= 3e-3, rel
    assert (out[~live] == 0).all()          # untouched (ops.cross_enc16 zero-fills the output)
    # everything finished: nothing runs, nothing is written
    out = ops.cross_enc16(qp, enc, T, heads, finished=torch.ones_like(fin), mode=mode)
    assert (out == 0).all()


def test_cross_enc_kernels_agree_on_decoder_like_inputs(cuda_ctx):
    """Both kernels on the same inputs at the bench geometry (TrOCR-base, 577 encoder tokens)."""
    from marie_icr_b200 import ops
    rows, T, heads, E = 160, 577, 16, 768
    qp, enc = _inputs(rows, T, heads, E, torch.float16, 21, qscale=0.08)
    a = ops.cross_enc16(qp, enc, T, heads, mode=0).float()
    b = ops.cross_enc16(qp, enc, T, heads, mode=1).float()
    assert ((a - b).norm() / b.norm()).item() <= 3e-3


@pytest.mark.parametrize("rows,beam,T,heads,E", [(4, 3, 577, 16, 768), (3, 5, 577, 16, 768), (3, 3, 577, 16, 1024),
                                                 (5, 2, 50, 2, 128), (2, 4, 130, 16, 768), (150, 3, 70, 16, 768),
                                                 (3, 8, 100, 16, 768), (2, 5, 65, 16, 1024), (7, 5, 33, 2, 128)])
def test_cross_enc_beams_share_the_pass(cuda_ctx, dtype16, rows, beam, T, heads, E):
    """beam >= 2: the hypotheses of a crop ([crop][beam] rows of qp) go through the tcgen05 kernel together — up to three
    per pass over the crop's encoder states (two for E = 1024), further groups as further passes."""
    from marie_icr_b200 import ops
    qp, enc = _inputs(rows, T, heads, E, dtype16, rows * T + beam, beam=beam)
    out = ops.cross_enc16(qp, enc, T, heads, mode=0, beam=beam).float()
    ref = _ref(qp, enc, rows, T, heads, beam)
    assert torch.isfinite(out).all()
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel <= (3e-3 if dtype16 == torch.float16 else 2e-2), rel
    worst = ((out - ref).view(rows * beam, -1).norm(dim=1) / ref.view(rows * beam, -1).norm(dim=1)).max().item()
    assert worst <= (6e-3 if dtype16 == torch.float16 else 4e-2), worst


def test_cross_enc_beams_skip_finished_crops(cuda_ctx):
    from marie_icr_b200 import ops
    rows, beam, T, heads, E = 200, 5, 70, 16, 768
    qp, enc = _inputs(rows, T, heads, E, torch.float16, 13, beam=beam)
    torch.manual_seed(6)
    fin = (torch.rand(rows, device="cuda") < 0.5).to(torch.uint8)
    out = ops.cross_enc16(qp, enc, T, heads, finished=fin, mode=0, beam=beam).float()
    ref = _ref(qp, enc, rows, T, heads, beam)
    live = (fin == 0).repeat_interleave(beam)
    assert ((out[live] - ref[live]).norm() / ref[live].norm()).item() <= 3e-3
    assert (out[~live] == 0).all()
