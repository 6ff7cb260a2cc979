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
