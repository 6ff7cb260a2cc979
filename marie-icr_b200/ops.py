"""Thin torch-tensor wrappers over the stage-level C-ABI entry points (used by tests and the plugins)."""
import ctypes

import numpy as np
import torch

from ._lib import Context, c_float, c_int, c_ll, c_void_p, cur_stream, ptr

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
OUT_BF16, OUT_F32, OUT_F32_PLANAR = 0, 1, 2


def _ctx(t):
    return Context.get(t.device.index or 0)


def gemm16(a, w, bias=None, act=ACT_NONE, residual=None, out_dtype=None, n_out=None):
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) (+ residual).  a, w 16-bit row-major; K % 64 == 0."""
    dt = _ctx(a).torch_dtype
    out_dtype = dt if out_dtype is None else out_dtype
    assert a.is_cuda and a.dtype == dt and w.dtype == dt, (a.dtype, w.dtype, dt)
    assert a.stride(-1) == 1 and w.is_contiguous()
    M, K = a.shape
    n_rows = w.shape[0]
    N = n_rows if n_out is None else n_out
    out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    _ctx(a).call(
        "mb_gemm16", ptr(a), c_ll(a.stride(0)), ptr(w), c_int(n_rows), c_int(M), c_int(N), c_int(K),
        ptr(bias), c_int(act), ptr(residual), c_ll(residual.stride(0) if residual is not None else 0),
        ptr(out), c_ll(out.stride(0)), c_int(OUT_F32 if out_dtype == torch.float32 else OUT_BF16),
        cur_stream())
    return out


def conv16(x0, w, bias=None, act=ACT_NONE, x1=None, taps=9, dil=1, n_out=None, out_dtype=None, planar=False):
    """NHWC implicit-GEMM convolution. x0/x1: [N,H,W,C] 16-bit; w: [rows, taps*(C0+C1)] 16-bit."""
    dt = _ctx(x0).torch_dtype
    out_dtype = dt if out_dtype is None else out_dtype
    assert x0.is_cuda and x0.dtype == dt and w.dtype == dt and x0.is_contiguous()
    n, h, wd, c0 = x0.shape
    c1 = 0
    if x1 is not None:
        assert x1.shape[:3] == x0.shape[:3] and x1.is_contiguous()
        c1 = x1.shape[3]
    rows = w.shape[0]
    N = rows if n_out is None else n_out
    if planar:
        out = torch.empty((N, n, h, wd), device=x0.device, dtype=torch.float32)   # channel planes over the batch
        mode, out_ld, plane = OUT_F32_PLANAR, 1, n * h * wd
    else:
        out = torch.empty((n, h, wd, N), device=x0.device, dtype=out_dtype)
        mode, out_ld, plane = (OUT_F32 if out_dtype == torch.float32 else OUT_BF16), N, 0
    _ctx(x0).call(
        "mb_conv16", ptr(x0), c_int(c0), c_int(c0), ptr(x1), c_int(c1), c_int(c1), c_int(n), c_int(h),
        c_int(wd), c_int(taps), c_int(dil), ptr(w), c_int(rows), c_int(N), ptr(bias), c_int(act), ptr(out),
        c_ll(out_ld), c_int(mode), c_ll(plane), cur_stream())
    return out


def pack_conv_weight(w_oihw, dtype=torch.float16):
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] 16-bit (k = (ky*3+kx)*Cin + c), the layout mb_conv16 reads."""
    co, ci, kh, kw = w_oihw.shape
    return w_oihw.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous().to(dtype)


def craft_post(text, link, text_threshold, link_threshold, low_text, ratios=None, page_hw=None, max_labels=8192,
               max_boxes=4096):
    """Batched score-map post-processing on device. text/link: [n,h,w] fp32 cuda.  Returns a dict of tensors."""
    assert text.is_cuda and text.dtype == torch.float32 and text.is_contiguous() and link.is_contiguous()
    n, h, w = text.shape
    dev = text.device
    out = dict(
        labels=torch.empty((n, h, w), dtype=torch.int32, device=dev),
        n_labels=torch.zeros((n,), dtype=torch.int32, device=dev),
        stats=torch.zeros((n, max_labels, 5), dtype=torch.int32, device=dev),
        det=torch.empty((n, max_boxes, 4, 2), dtype=torch.float32, device=dev),
        adj=torch.empty((n, max_boxes, 4, 2), dtype=torch.float32, device=dev),
        rects=torch.empty((n, max_boxes, 4), dtype=torch.int32, device=dev),
        mapper=torch.empty((n, max_boxes), dtype=torch.int32, device=dev),
        n_boxes=torch.zeros((n,), dtype=torch.int32, device=dev),
    )
    if ratios is not None:
        ratios = torch.as_tensor(ratios, dtype=torch.float64).reshape(n, 2).to(dev)
    if page_hw is not None:
        page_hw = torch.as_tensor(page_hw, dtype=torch.int32).reshape(n, 2).to(dev)
    _ctx(text).call(
        "mb_craft_post", ptr(text), ptr(link), c_int(n), c_int(h), c_int(w), c_float(text_threshold),
        c_float(link_threshold), c_float(low_text), ptr(ratios), ptr(page_hw), ptr(out["labels"]),
        ptr(out["n_labels"]), ptr(out["stats"]), c_int(max_labels), ptr(out["det"]), ptr(out["adj"]),
        ptr(out["rects"]), ptr(out["mapper"]), ptr(out["n_boxes"]), c_int(max_boxes), cur_stream())
    return out


def craft_canvas_dims(page_h, page_w, canvas_size=None, mag_ratio=1.0):
    """Size arithmetic of resize_aspect_ratio (marie/models/craft/imgproc.py:45-70): returns
    (target_h, target_w, out_h, out_w, ratio)."""
    canvas_size = page_w if canvas_size is None else canvas_size
    target_size = mag_ratio * max(page_h, page_w)
    if target_size > canvas_size:
        target_size = canvas_size
    ratio = target_size / max(page_h, page_w)
    th, tw = int(page_h * ratio), int(page_w * ratio)
    oh = th if th % 32 == 0 else th + (32 - th % 32)
    ow = tw if tw % 32 == 0 else tw + (32 - tw % 32)
    return th, tw, oh, ow, ratio


def page_preprocess(pages_u8, canvas_size=None, mag_ratio=1.0):
    """pages_u8: [n, H, W, 3] u8 cuda (BGR). Returns (x [n, oh, ow, 4] bf16 NHWC, ratio)."""
    assert pages_u8.is_cuda and pages_u8.dtype == torch.uint8 and pages_u8.is_contiguous()
    n, ph, pw, _ = pages_u8.shape
    th, tw, oh, ow, ratio = craft_canvas_dims(ph, pw, canvas_size, mag_ratio)
    out = torch.empty((n, oh, ow, 4), dtype=_ctx(pages_u8).torch_dtype, device=pages_u8.device)
    _ctx(pages_u8).call("mb_page_preprocess", ptr(pages_u8), c_int(n), c_int(ph), c_int(pw), c_int(th), c_int(tw),
                        c_int(oh), c_int(ow), ptr(out), cur_stream())
    return out, ratio


def pack_crops(pages_u8, rects, page_idx, layout=0):
    """rects [n,4] i32 (x,y,w,h), page_idx [n] i32 on device -> [n,3,384,384] or [n*576,768] bf16."""
    n = rects.shape[0]
    _, ph, pw, _ = pages_u8.shape
    shape = (n, 3, 384, 384) if layout == 0 else (n * 576, 768)
    out = torch.empty(shape, dtype=_ctx(pages_u8).torch_dtype, device=pages_u8.device)
    _ctx(pages_u8).call("mb_pack_crops", ptr(pages_u8), c_int(ph), c_int(pw), ptr(rects.contiguous()),
                        ptr(page_idx.contiguous()), c_int(n), ptr(out), c_int(layout), cur_stream())
    return out


def pack_fragments(fragments, device="cuda", layout=0):
    """fragments: list of [h,w,3] u8 BGR numpy arrays (host).  One H2D copy of the packed bytes."""
    import numpy as np
    n = len(fragments)
    sizes = [int(f.shape[0]) * int(f.shape[1]) * 3 for f in fragments]
    offsets = np.zeros(n, np.int64)
    if n > 1:
        offsets[1:] = np.cumsum(np.asarray(sizes[:-1], np.int64))
    buf = np.empty(int(sum(sizes)), np.uint8)
    for f, o, s in zip(fragments, offsets, sizes):
        buf[o:o + s] = np.ascontiguousarray(f, dtype=np.uint8).reshape(-1)
    hw = np.array([[f.shape[0], f.shape[1]] for f in fragments], np.int32).reshape(n, 2)
    dbuf = torch.from_numpy(buf).to(device)
    doff = torch.from_numpy(offsets).to(device)
    dhw = torch.from_numpy(hw).to(device)
    shape = (n, 3, 384, 384) if layout == 0 else (n * 576, 768)
    out = torch.empty(shape, dtype=_ctx(dbuf).torch_dtype, device=device)
    _ctx(dbuf).call("mb_pack_fragments", ptr(dbuf), ptr(doff), ptr(dhw), c_int(n), ptr(out), c_int(layout),
                    cur_stream())
    return out


def load_craft(blob: bytes, device=0):
    buf = ctypes.create_string_buffer(blob, len(blob))
    Context.get(device).call("mb_load_craft", buf, ctypes.c_size_t(len(blob)))


def craft_forward(x, want_feature=False):
    """x: [n,h,w,4] 16-bit NHWC (page_preprocess output). Returns scores [2,n,h/2,w/2] fp32 (text, link)
    and optionally feature [n,h/2,w/2,64]."""
    assert x.is_cuda and x.dtype == _ctx(x).torch_dtype and x.is_contiguous() and x.shape[3] == 4
    n, h, w, _ = x.shape
    scores = torch.empty((2, n, h // 2, w // 2), dtype=torch.float32, device=x.device)
    feat = torch.empty((n, h // 2, w // 2, 64), dtype=x.dtype, device=x.device) if want_feature else None
    _ctx(x).call("mb_craft_forward", ptr(x), c_int(n), c_int(h), c_int(w), ptr(scores), ptr(feat), cur_stream())
    return (scores, feat) if want_feature else scores


def load_refine(blob: bytes, device=0):
    buf = ctypes.create_string_buffer(blob, len(blob))
    Context.get(device).call("mb_load_refine", buf, ctypes.c_size_t(len(blob)))


def refine_forward(scores, feature):
    """RefineNet.forward on the device: scores [2,n,h,w] fp32 + feature [n,h,w,64] 16-bit (craft_forward(want_feature=True);
    its channels 32/33 are overwritten) -> refined link map [n,h,w] fp32."""
    assert scores.is_cuda and scores.dtype == torch.float32 and scores.is_contiguous() and feature.is_contiguous()
    _, n, h, w = scores.shape
    assert tuple(feature.shape) == (n, h, w, 64) and feature.dtype == _ctx(scores).torch_dtype
    out = torch.empty((n, h, w), dtype=torch.float32, device=scores.device)
    _ctx(scores).call("mb_refine_forward", ptr(feature), ptr(scores), c_int(n), c_int(h), c_int(w), ptr(out), cur_stream())
    return out


def line_components(link, link_threshold, max_labels=8192, want_labels=False):
    """Line branch of get_prediction up to the component list: link [n,h,w] fp32 -> dict(n_labels [n], stats
    [n,max_labels,5] (left, top, width, height, area), labels [n,h,w] when asked)."""
    assert link.is_cuda and link.dtype == torch.float32 and link.is_contiguous()
    n, h, w = link.shape
    dev = link.device
    out = dict(n_labels=torch.zeros((n,), dtype=torch.int32, device=dev),
               stats=torch.empty((n, max_labels, 5), dtype=torch.int32, device=dev),
               labels=torch.empty((n, h, w), dtype=torch.int32, device=dev) if want_labels else None)
    _ctx(link).call("mb_line_components", ptr(link), c_int(n), c_int(h), c_int(w), c_float(link_threshold),
                    ptr(out["labels"]), ptr(out["n_labels"]), ptr(out["stats"]), c_int(max_labels), cur_stream())
    return out


def line_boxes(link, link_threshold, ratio_w, ratio_h, ratio_net=2, max_labels=8192):
    """Refiner line boxes in page coordinates, per image: components -> line_merge (host) -> int() scaling
    (marie/boxes/craft_box_processor.py:161-217)."""
    from . import lines as _lines
    comp = line_components(link, link_threshold, max_labels)
    nl = comp["n_labels"].cpu().tolist()
    stats = comp["stats"].cpu().numpy()
    res = []
    for i, k in enumerate(nl):
        boxes = stats[i, 1:k, :4].tolist()
        merged = _lines.line_merge(boxes) if boxes else []
        res.append([[int(b[0] * ratio_w * ratio_net), int(b[1] * ratio_h * ratio_net), int(b[2] * ratio_w * ratio_net),
                     int(b[3] * ratio_h * ratio_net)] for b in np.asarray(merged).tolist()])
    return res


# ------------------------------------------------------------------------------------------------ TrOCR
def load_trocr(blob: bytes, device=0):
    buf = ctypes.create_string_buffer(blob, len(blob))
    Context.get(device).call("mb_load_trocr", buf, ctypes.c_size_t(len(blob)))


def trocr_dims(device=0):
    d = (ctypes.c_int * 4)()
    Context.get(device).call("mb_trocr_dims", d)
    return dict(enc_dim=d[0], dec_dim=d[1], vocab=d[2], tokens=d[3])


def trocr_stats(device=0):
    d = (ctypes.c_ulonglong * 3)()
    Context.get(device).call("mb_trocr_stats", d)
    return dict(decode_calls=int(d[0]), decode_steps=int(d[1]), decode_rows=int(d[2]))


def trocr_encode(patches, out=None):
    """patches [n*576, 768] 16-bit (pack_crops / pack_fragments layout=1) -> enc_out [n, 577, enc_dim] (written into
    `out` when given: a contiguous [n, 577, enc_dim] view, e.g. a slice of a larger decode batch)."""
    ctx = _ctx(patches)
    dims = trocr_dims(ctx.device)
    n = patches.shape[0] // (dims["tokens"] - 1)
    if out is None:
        out = torch.empty((n, dims["tokens"], dims["enc_dim"]), dtype=patches.dtype, device=patches.device)
    assert out.is_contiguous() and tuple(out.shape) == (n, dims["tokens"], dims["enc_dim"]) and out.dtype == patches.dtype
    ctx.call("mb_trocr_encode", ptr(patches), c_int(n), ptr(out), cur_stream())
    return out


def trocr_decode(enc_out, beam=1, max_len_b=200, out_ld=None):
    """enc_out [n, 577, D] -> (tokens [n, out_ld] i32 incl. final EOS, lengths [n] i32, scores [n] f32, steps_run)."""
    ctx = _ctx(enc_out)
    n = enc_out.shape[0]
    out_ld = (max_len_b + 1) if out_ld is None else out_ld
    tokens = torch.empty((n, out_ld), dtype=torch.int32, device=enc_out.device)
    lengths = torch.empty((n,), dtype=torch.int32, device=enc_out.device)
    scores = torch.empty((n,), dtype=torch.float32, device=enc_out.device)
    steps = ctypes.c_int(0)
    ctx.call("mb_trocr_decode", ptr(enc_out.contiguous()), c_int(n), c_int(beam), c_int(max_len_b), ptr(tokens),
             c_int(out_ld), ptr(lengths), ptr(scores), ctypes.byref(steps), cur_stream())
    return tokens, lengths, scores, steps.value


def trocr_forced_logits(enc_out, forced):
    """Teacher-forced decoder logits: forced [n, L] i32 -> [L, n, vocab] fp32."""
    ctx = _ctx(enc_out)
    n, L = forced.shape
    vocab = trocr_dims(ctx.device)["vocab"]
    out = torch.empty((L, n, vocab), dtype=torch.float32, device=enc_out.device)
    ctx.call("mb_trocr_forced_logits", ptr(enc_out.contiguous()), c_int(n), ptr(forced.contiguous()), c_int(L), ptr(out),
             cur_stream())
    return out


def trocr_recognize(patches, beam=1, max_len_b=200, chunk=0, out_ld=None):
    ctx = _ctx(patches)
    n = patches.shape[0] // (trocr_dims(ctx.device)["tokens"] - 1)
    out_ld = (max_len_b + 1) if out_ld is None else out_ld
    tokens = torch.empty((n, out_ld), dtype=torch.int32, device=patches.device)
    lengths = torch.empty((n,), dtype=torch.int32, device=patches.device)
    scores = torch.empty((n,), dtype=torch.float32, device=patches.device)
    ctx.call("mb_trocr_recognize", ptr(patches), c_int(n), c_int(beam), c_int(max_len_b), c_int(chunk), ptr(tokens),
             c_int(out_ld), ptr(lengths), ptr(scores), cur_stream())
    return tokens, lengths, scores


def attention16(qkv, n, T, scale=0.125, mode=0):
    """qkv [n*T, 3*D] 16-bit -> softmax(Q K^T * scale) V [n*T, D].  mode 0: tcgen05 kernel, 1: mma.sync kernel."""
    D = qkv.shape[1] // 3
    out = torch.empty((n * T, D), dtype=qkv.dtype, device=qkv.device)
    _ctx(qkv).call("mb_attention16", ptr(qkv.contiguous()), ptr(out), c_int(n), c_int(T), c_int(D), c_float(scale),
                   c_int(mode), cur_stream())
    return out


def gemm16_res_stats(a, w, bias, residual, eps=1e-6):
    """out = a @ w.T + bias + residual (16-bit) and the LayerNorm statistics (-mean, rstd) [M, 2] fp32 of its rows,
    produced by the GEMM's own epilogue (mb_gemm16_res_stats)."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    part = torch.empty((M * 4 * ((N + 255) // 256),), dtype=torch.float32, device=a.device)
    stats = torch.empty((M, 2), dtype=torch.float32, device=a.device)
    _ctx(a).call("mb_gemm16_res_stats", ptr(a.contiguous()), ptr(w.contiguous()), ptr(bias), ptr(residual.contiguous()), ptr(out),
                 c_ll(M), c_int(N), c_int(K), c_float(eps), ptr(part), ptr(stats), cur_stream())
    return out, stats


def cross_enc16(qp, enc, T, heads, finished=None, mode=0, beam=1):
    """Cross-attention core over the encoder states.  qp [crops*beam, heads*E] ([crop][beam] rows), enc [crops*T, E]
    16-bit -> [crops*beam, heads*E].  mode 0: tcgen05 / TMA kernel (the hypotheses of a crop share the pass), 1: mma.sync
    kernel (beam 1 only).  finished: optional [crops] uint8 mask of crops to skip."""
    rows = qp.shape[0] // beam
    E = enc.shape[1]
    out = torch.zeros_like(qp)
    live = torch.empty((rows + 1,), dtype=torch.int32, device=qp.device)
    _ctx(qp).call("mb_cross_enc16", ptr(qp.contiguous()), ptr(enc.contiguous()), ptr(out), c_int(rows), c_int(beam), c_int(T),
                  c_int(heads), c_int(E), ptr(finished), ptr(live), c_int(mode), cur_stream())
    return out


def gemm16_batched(a, w, batches, n, k, a_col_stride, w_row_stride, out_col_stride, out_cols, bias=None, act=ACT_NONE):
    """Block-diagonal GEMM (see mb_gemm16_batched).  a [M, lda], w [rows, k] -> out [M, out_cols]."""
    M = a.shape[0]
    out = torch.zeros((M, out_cols), device=a.device, dtype=a.dtype)
    _ctx(a).call("mb_gemm16_batched", ptr(a), c_ll(a.stride(0)), ptr(w), c_int(w.shape[0]), c_int(M), c_int(n), c_int(k),
                 c_int(batches), c_int(a_col_stride), c_int(w_row_stride), c_int(out_col_stride), ptr(bias), c_int(act),
                 ptr(out), c_ll(out.stride(0)), cur_stream())
    return out


def layernorm16(x, gamma, beta, eps=1e-6):
    out = torch.empty_like(x)
    rows, D = x.shape
    _ctx(x).call("mb_layernorm16", ptr(x), ptr(out), ptr(gamma), ptr(beta), c_ll(rows), c_int(D), c_float(eps), cur_stream())
    return out
