"""Weight packers: PyTorch state dicts -> the flat blobs read by mb_load_craft / mb_load_trocr (csrc/blob.cuh).

CRAFT   reference module: marie/models/craft/craft.py:31-57 + basenet/vgg16_bn.py:23-47 (state-dict keys below are
        the reference's own, after copyStateDict strips `module.`, marie/boxes/box_processor.py:15-24).
        BatchNorm (eval) is folded into the preceding convolution in fp32, then weights are rounded to bf16 once.
TrOCR   reference module: TrOCRModel (marie/models/unilm/trocr/trocr_models.py:126-147,492-524): timm ViT encoder
        (`encoder.deit.*`) + fairseq TransformerDecoder (`decoder.*`).
"""
import struct

import numpy as np
import torch

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.int32: 2, torch.float16: 3}


def build_blob(tensors):
    """tensors: dict name -> torch tensor (f32 / bf16 / i32, CPU). Returns bytes."""
    names = list(tensors)
    entry_bytes = 64 + 4 + 4 + 32 + 8 + 8
    header = 16 + len(names) * entry_bytes
    data_off = (header + 255) // 256 * 256
    entries, chunks, off = [], [], 0
    for name in names:
        t = tensors[name].detach().cpu().contiguous()
        assert t.dtype in _DT, (name, t.dtype)
        raw = t.view(torch.int16).numpy().tobytes() if t.dtype in (torch.bfloat16, torch.float16) else t.numpy().tobytes()
        dims = list(t.shape) + [0] * (4 - t.dim())
        assert t.dim() <= 4 and len(name) < 64
        entries.append(name.encode().ljust(64, b"\0") + struct.pack("<II4QQQ", _DT[t.dtype], t.dim(), *dims, off, len(raw)))
        pad = (-len(raw)) % 256
        chunks.append(raw + b"\0" * pad)
        off += len(raw) + pad
    blob = b"MB2W" + struct.pack("<III", 1, len(names), 0) + b"".join(entries)
    blob += b"\0" * (data_off - len(blob))
    return blob + b"".join(chunks)


def read_blob(blob):
    """Inverse of build_blob (tests / debugging): blob bytes -> dict name -> torch tensor."""
    assert blob[:4] == b"MB2W"
    _, count, _ = struct.unpack_from("<III", blob, 4)
    entry_bytes = 64 + 4 + 4 + 32 + 8 + 8
    data_off = (16 + count * entry_bytes + 255) // 256 * 256
    inv = {v: k for k, v in _DT.items()}
    out = {}
    for i in range(count):
        base = 16 + i * entry_bytes
        name = blob[base:base + 64].split(b"\0", 1)[0].decode()
        dt, nd, d0, d1, d2, d3, off, nbytes = struct.unpack_from("<II4QQQ", blob, base + 64)
        dtype = inv[dt]
        raw = np.frombuffer(blob, dtype=np.uint8, count=nbytes, offset=data_off + off).copy()
        if dtype in (torch.bfloat16, torch.float16):
            t = torch.from_numpy(raw.view(np.int16)).view(dtype)
        elif dtype == torch.int32:
            t = torch.from_numpy(raw.view(np.int32))
        else:
            t = torch.from_numpy(raw.view(np.float32))
        out[name] = t.reshape([d0, d1, d2, d3][:nd])
    return out


def _fold_bn(w, b, sd, bn_key, eps=1e-5):
    g, beta = sd[bn_key + ".weight"].float(), sd[bn_key + ".bias"].float()
    mu, var = sd[bn_key + ".running_mean"].float(), sd[bn_key + ".running_var"].float()
    s = g / torch.sqrt(var + eps)
    return w * s[:, None, None, None], beta + (b - mu) * s


def _pack_conv(w_oihw, rows=None, cin_pad=None, dtype=torch.float16):
    """[Cout,Cin,kh,kw] f32 -> [rows, kh*kw*cin_pad] bf16, k = (ky*kw+kx)*cin_pad + c (zero padding)."""
    co, ci, kh, kw = w_oihw.shape
    cin_pad = ci if cin_pad is None else cin_pad
    rows = co if rows is None else rows
    out = torch.zeros(rows, kh * kw, cin_pad, dtype=torch.float32)
    out[:co, :, :ci] = w_oihw.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
    return out.reshape(rows, kh * kw * cin_pad).to(dtype)


def _pad_bias(b, rows):
    out = torch.zeros(rows, dtype=torch.float32)
    out[: b.numel()] = b.float()
    return out


# (blob layer name, conv key, bn key or None, rows, cin_pad)
_CRAFT_LAYERS = [
    ("conv1_2", "basenet.slice1.3", "basenet.slice1.4", None, None),
    ("conv2_1", "basenet.slice1.7", "basenet.slice1.8", None, None),
    ("conv2_2", "basenet.slice1.10", "basenet.slice1.11", None, None),
    ("conv3_1", "basenet.slice2.14", "basenet.slice2.15", None, None),
    ("conv3_2", "basenet.slice2.17", "basenet.slice2.18", None, None),
    ("conv3_3", "basenet.slice3.20", "basenet.slice3.21", None, None),
    ("conv4_1", "basenet.slice3.24", "basenet.slice3.25", None, None),
    ("conv4_2", "basenet.slice3.27", "basenet.slice3.28", None, None),
    ("conv4_3", "basenet.slice4.30", "basenet.slice4.31", None, None),
    ("conv5_1", "basenet.slice4.34", "basenet.slice4.35", None, None),
    ("conv5_2", "basenet.slice4.37", "basenet.slice4.38", None, None),
    ("fc6", "basenet.slice5.1", None, None, None),
    ("fc7", "basenet.slice5.2", None, None, None),
    ("upconv1a", "upconv1.conv.0", "upconv1.conv.1", None, None),
    ("upconv1b", "upconv1.conv.3", "upconv1.conv.4", None, None),
    ("upconv2a", "upconv2.conv.0", "upconv2.conv.1", None, None),
    ("upconv2b", "upconv2.conv.3", "upconv2.conv.4", None, None),
    ("upconv3a", "upconv3.conv.0", "upconv3.conv.1", None, None),
    ("upconv3b", "upconv3.conv.3", "upconv3.conv.4", None, None),
    ("upconv4a", "upconv4.conv.0", "upconv4.conv.1", None, None),
    ("upconv4b", "upconv4.conv.3", "upconv4.conv.4", 64, None),   # 32 real channels stored padded to 64
    ("cls1", "conv_cls.0", None, 64, 64),
    ("cls2", "conv_cls.2", None, 64, 64),
    ("cls3", "conv_cls.4", None, 64, 64),
    ("cls4", "conv_cls.6", None, 64, 64),
    ("cls5", "conv_cls.8", None, 16, 64),
]


def pack_craft(state_dict, dtype=torch.float16):
    """Reference CRAFT state dict -> blob bytes for a context whose element type is `dtype` (fp16 or bf16)."""
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state_dict.items()}
    t = {}
    w, b = _fold_bn(sd["basenet.slice1.0.weight"].float(), sd["basenet.slice1.0.bias"].float(), sd, "basenet.slice1.1")
    w = w.to(dtype).float()                               # rounded once, like every other layer
    t["conv1_1.w"] = w.permute(2, 3, 1, 0).reshape(27, 64).contiguous()   # [(ky*3+kx)*3+c][cout]
    t["conv1_1.b"] = b.contiguous()
    for name, ck, bk, rows, cin_pad in _CRAFT_LAYERS:
        w, b = sd[ck + ".weight"].float(), sd[ck + ".bias"].float()
        if bk is not None:
            w, b = _fold_bn(w, b, sd, bk)
        rows_ = w.shape[0] if rows is None else rows
        t[name + ".w"] = _pack_conv(w, rows_, cin_pad, dtype)
        t[name + ".b"] = _pad_bias(b, rows_)
    return build_blob(t)


def pack_refine(state_dict, dtype=torch.float16):
    """Reference RefineNet state dict (marie/models/craft/refinenet.py:15-55) -> blob bytes for mb_load_refine.
    BatchNorm folded in fp32, weights rounded once.  The first layer's 34 input channels (text, link, feature[0:32]) are
    permuted to the device layout (feature[0:32], text, link, zero padding to 64); the four final 1x1 -> 1 convolutions
    become one [16, 512] matrix over the four branches' hidden states laid side by side (their biases add up)."""
    sd = {k[len("module."):] if k.startswith("module.") else k: v.detach().float().cpu() for k, v in state_dict.items()
          if torch.is_tensor(v)}
    t = {}
    for i, (ci, bi) in enumerate(((0, 1), (3, 4), (6, 7)), 1):
        w, b = _fold_bn(sd[f"last_conv.{ci}.weight"], sd[f"last_conv.{ci}.bias"], sd, f"last_conv.{bi}")
        if i == 1:
            w = torch.cat([w[:, 2:34], w[:, 0:2]], 1)
        t[f"ref.c{i}.w"] = _pack_conv(w, 64, 64, dtype)
        t[f"ref.c{i}.b"] = _pad_bias(b, 64)
    fin_w = torch.zeros(16, 512)
    fin_b = torch.zeros(16)
    for k in range(1, 5):
        w, b = _fold_bn(sd[f"aspp{k}.0.weight"], sd[f"aspp{k}.0.bias"], sd, f"aspp{k}.1")
        t[f"ref.a{k}a.w"], t[f"ref.a{k}a.b"] = _pack_conv(w, 128, 64, dtype), _pad_bias(b, 128)
        w, b = _fold_bn(sd[f"aspp{k}.3.weight"], sd[f"aspp{k}.3.bias"], sd, f"aspp{k}.4")
        t[f"ref.a{k}b.w"], t[f"ref.a{k}b.b"] = _pack_conv(w, 128, 128, dtype), _pad_bias(b, 128)
        fin_w[0, (k - 1) * 128:k * 128] = sd[f"aspp{k}.6.weight"].reshape(128)
        fin_b[0] += sd[f"aspp{k}.6.bias"].reshape(())
    t["ref.final.w"], t["ref.final.b"] = fin_w.to(dtype), fin_b
    return build_blob(t)


def validate_trocr_variant(state_dict, cfg, info=None):
    """The kernels implement ONE TrOCR variant — the reference's default checkpoint family (trocr-{base,large}-printed /
    -handwritten, arch tables trocr_models.py:423-447): BEiT encoder without qkv bias or distillation token, 577 tokens,
    sinusoidal decoder positions, scaled embeddings, no embedding LayerNorm, post-LN decoder layers, ReLU.  Any other
    variant the reference can load (trocr_small with learned positions / layernorm_embedding, deit_* encoders with qkv
    bias and dist_token, GELU decoders) would silently produce wrong tokens — refuse it instead.  `info`: configuration
    fields read from the checkpoint (checkpoint.load_fairseq_checkpoint), when there are any."""
    bad = []
    keys = set(state_dict.keys())
    if "decoder.embed_positions.weight" in keys:
        bad.append("learned decoder positions (decoder.embed_positions.weight); only sinusoidal positions are implemented")
    if any(k.startswith("decoder.layernorm_embedding.") for k in keys):
        bad.append("decoder.layernorm_embedding")
    if any(k.startswith("decoder.layer_norm.") for k in keys):
        bad.append("decoder.layer_norm (pre-LN decoder, decoder_normalize_before)")
    if any(k.endswith("attn.qkv.bias") or k.endswith("attn.q_bias") for k in keys if k.startswith("encoder.deit.")):
        bad.append("encoder qkv bias")
    if "encoder.deit.dist_token" in keys:
        bad.append("encoder.deit.dist_token (distilled DeiT encoder)")
    if getattr(cfg, "tokens", 577) != 577:
        bad.append(f"{cfg.tokens} encoder tokens (577 expected: 384x384 input, 16x16 patches, cls token)")
    if "encoder.deit.pos_embed" in keys and tuple(state_dict["encoder.deit.pos_embed"].shape[:2]) != (1, 577):
        bad.append(f"pos_embed of shape {tuple(state_dict['encoder.deit.pos_embed'].shape)}")
    if "decoder.output_projection.weight" not in keys:
        bad.append("no decoder.output_projection.weight (tied embeddings are not implemented)")
    for k in keys:
        if k.startswith("encoder.deit.blocks.") and (".gamma_1" in k or ".gamma_2" in k or "relative_position" in k):
            bad.append("BEiT LayerScale / relative position bias (" + k + ")")
            break
    info = info or {}
    act = info.get("activation_fn")
    if act is not None and str(act).split(".")[-1].lower() != "relu":
        bad.append(f"decoder activation_fn={act!r} (ReLU is implemented)")
    for flag, why in (("decoder_learned_pos", "learned positions"), ("decoder_normalize_before", "pre-LN decoder"),
                      ("layernorm_embedding", "embedding LayerNorm"), ("no_scale_embedding", "unscaled embeddings")):
        if info.get(flag) is True:
            bad.append(f"{flag}=True ({why})")
    if bad:
        raise ValueError("unsupported TrOCR variant: " + "; ".join(bad))


def pack_trocr(state_dict, cfg, dtype=torch.float16, info=None):
    """fairseq TrOCR state dict (`encoder.deit.*`, `decoder.*`) -> blob bytes for mb_load_trocr.
    cfg: any object with enc_dim, enc_layers, enc_heads, enc_ffn, dec_dim, dec_layers, dec_heads, dec_ffn, vocab,
    tokens, max_positions.  Layout choices: q/k/v of the decoder self-attention fused into one [3H, H] matrix,
    cross k/v into [2H, enc_dim]; the fairseq query scaling head_dim^-0.5 = 0.125 (a power of two, exact) is folded
    into the decoder q weights/biases; cls_token is folded into row 0 of the position table; the sinusoidal position
    table (fairseq SinusoidalPositionalEmbedding) is precomputed in fp32."""
    import math
    validate_trocr_variant(state_dict, cfg, info)
    sd = {k: v.detach().float().cpu() for k, v in state_dict.items() if torch.is_tensor(v)}
    t = {}
    t["config"] = torch.tensor([cfg.enc_dim, cfg.enc_layers, cfg.enc_heads, cfg.enc_ffn, cfg.dec_dim, cfg.dec_layers,
                                cfg.dec_heads, cfg.dec_ffn, cfg.vocab, cfg.tokens, cfg.max_positions], dtype=torch.int32)
    e = "encoder.deit."
    D, H = cfg.enc_dim, cfg.dec_dim
    assert D // cfg.enc_heads == 64 and H // cfg.dec_heads == 64, "head dim must be 64"
    t["enc.patch.w"] = sd[e + "patch_embed.proj.weight"].reshape(D, -1).to(dtype)
    t["enc.patch.b"] = sd[e + "patch_embed.proj.bias"]
    cls_pos = sd[e + "pos_embed"][0].clone()
    cls_pos[0] += sd[e + "cls_token"][0, 0]
    t["enc.cls_pos"] = cls_pos
    t["enc.norm.w"], t["enc.norm.b"] = sd[e + "norm.weight"], sd[e + "norm.bias"]
    for i in range(cfg.enc_layers):
        b, p = f"{e}blocks.{i}.", f"enc.L{i}."
        t[p + "ln1.w"], t[p + "ln1.b"] = sd[b + "norm1.weight"], sd[b + "norm1.bias"]
        t[p + "ln2.w"], t[p + "ln2.b"] = sd[b + "norm2.weight"], sd[b + "norm2.bias"]
        t[p + "qkv.w"] = sd[b + "attn.qkv.weight"].to(dtype)
        t[p + "proj.w"], t[p + "proj.b"] = sd[b + "attn.proj.weight"].to(dtype), sd[b + "attn.proj.bias"]
        t[p + "fc1.w"], t[p + "fc1.b"] = sd[b + "mlp.fc1.weight"].to(dtype), sd[b + "mlp.fc1.bias"]
        t[p + "fc2.w"], t[p + "fc2.b"] = sd[b + "mlp.fc2.weight"].to(dtype), sd[b + "mlp.fc2.bias"]
        # pre-LN blocks: LN(x) W^T + b = rstd * (x (W * gamma)^T - mean * c) + (b + W beta), c = row sums of the ROUNDED
        # folded weights — the GEMM then reads the raw residual stream and the LayerNorm kernels disappear
        for name, ln, bias in (("qkv", "norm1", None), ("fc1", "norm2", sd[b + "mlp.fc1.bias"])):
            w = sd[b + ("attn.qkv.weight" if name == "qkv" else "mlp.fc1.weight")]
            wf = (w * sd[b + ln + ".weight"][None, :]).to(dtype)
            t[p + name + ".wf"] = wf
            t[p + name + ".c"] = wf.double().sum(1).float()
            bf = (w.double() @ sd[b + ln + ".bias"].double()).float()
            t[p + name + ".bf"] = bf if bias is None else bf + bias
    t["dec.embed"] = sd["decoder.embed_tokens.weight"].to(dtype)
    half = H // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    ang = torch.arange(cfg.max_positions + 2, dtype=torch.float)[:, None] * freq[None]
    pe = torch.cat([ang.sin(), ang.cos()], 1)
    pe[1] = 0                                               # padding_idx
    t["dec.pe"] = pe.contiguous()
    t["dec.out.w"] = sd["decoder.output_projection.weight"].to(dtype)
    sc = 64 ** -0.5
    for i in range(cfg.dec_layers):
        b, p = f"decoder.layers.{i}.", f"dec.L{i}."
        sa, ca = b + "self_attn.", b + "encoder_attn."
        t[p + "self.qkv.w"] = torch.cat([sd[sa + "q_proj.weight"] * sc, sd[sa + "k_proj.weight"], sd[sa + "v_proj.weight"]], 0).to(dtype)
        t[p + "self.qkv.b"] = torch.cat([sd[sa + "q_proj.bias"] * sc, sd[sa + "k_proj.bias"], sd[sa + "v_proj.bias"]], 0)
        t[p + "self.out.w"], t[p + "self.out.b"] = sd[sa + "out_proj.weight"].to(dtype), sd[sa + "out_proj.bias"]
        t[p + "cross.q.w"], t[p + "cross.q.b"] = (sd[ca + "q_proj.weight"] * sc).to(dtype), sd[ca + "q_proj.bias"] * sc
        t[p + "cross.kv.w"] = torch.cat([sd[ca + "k_proj.weight"], sd[ca + "v_proj.weight"]], 0).to(dtype)
        t[p + "cross.kv.b"] = torch.cat([sd[ca + "k_proj.bias"], sd[ca + "v_proj.bias"]], 0)
        # per-head transposed key projection for the cache-free greedy cross-attention: row (h, j) = Wk[h*64:(h+1)*64, j]
        wk = sd[ca + "k_proj.weight"]
        t[p + "cross.kT.w"] = wk.view(H // 64, 64, wk.shape[1]).transpose(1, 2).reshape(-1, 64).to(dtype)
        t[p + "cross.out.w"], t[p + "cross.out.b"] = sd[ca + "out_proj.weight"].to(dtype), sd[ca + "out_proj.bias"]
        t[p + "fc1.w"], t[p + "fc1.b"] = sd[b + "fc1.weight"].to(dtype), sd[b + "fc1.bias"]
        t[p + "fc2.w"], t[p + "fc2.b"] = sd[b + "fc2.weight"].to(dtype), sd[b + "fc2.bias"]
        for j, name in enumerate(("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"), 1):
            t[p + f"ln{j}.w"], t[p + f"ln{j}.b"] = sd[b + name + ".weight"], sd[b + name + ".bias"]
    return build_blob({k: v.contiguous() for k, v in t.items()})
