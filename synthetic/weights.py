"""Seeded synthetic weights for tests and bench.py — neither product code nor oracle code.

The reference ships no weights (model_zoo/ holds only a CLIP stub) and there is no network, so every run uses
random-init weights of the reference's architectures, rounded once so that the CPU oracle and the device see the same
values.  Two calibrations make the synthetic models behave like trained ones where the path needs it:
the CRAFT 'glyph path' (text-like score maps -> ~one box per word) and the TrOCR EOS row (hypotheses end after a few
tokens instead of running to max_len = 200).
"""
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

# (conv index, bn index) pairs per slice of torchvision vgg16_bn.features as cut by vgg16_bn.py:33-40
_VGG = {
    "slice1": [(0, 1), (3, 4), (7, 8), (10, 11)],
    "slice2": [(14, 15), (17, 18)],
    "slice3": [(20, 21), (24, 25), (27, 28)],
    "slice4": [(30, 31), (34, 35), (37, 38)],
}
_VGG_CH = {0: (3, 64), 3: (64, 64), 7: (64, 128), 10: (128, 128), 14: (128, 256), 17: (256, 256), 20: (256, 256),
           24: (256, 512), 27: (512, 512), 30: (512, 512), 34: (512, 512), 37: (512, 512)}
_UP = {"upconv1": (1024, 512, 256), "upconv2": (512, 256, 128), "upconv3": (256, 128, 64), "upconv4": (128, 64, 32)}
_CLS = {0: (32, 32, 3), 2: (32, 32, 3), 4: (32, 16, 3), 6: (16, 16, 1), 8: (16, 2, 1)}


def synth_craft_state(seed=0, random_bn=False, bf16_round=True):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cin, cout, k):
        w = torch.empty(cout, cin, k, k)
        fan_in, fan_out = cin * k * k, cout * k * k
        bound = (6.0 / (fan_in + fan_out)) ** 0.5            # xavier_uniform_, gain 1
        w.uniform_(-bound, bound, generator=g)
        if bf16_round:
            w = w.to(torch.bfloat16).float()
        sd[key + ".weight"] = w
        sd[key + ".bias"] = torch.zeros(cout)

    def bn(key, c):
        if random_bn:
            sd[key + ".weight"] = torch.empty(c).uniform_(0.6, 1.4, generator=g)
            sd[key + ".bias"] = torch.empty(c).uniform_(-0.2, 0.2, generator=g)
            sd[key + ".running_mean"] = torch.empty(c).uniform_(-0.2, 0.2, generator=g)
            sd[key + ".running_var"] = torch.empty(c).uniform_(0.5, 1.5, generator=g)
        else:
            sd[key + ".weight"] = torch.ones(c)
            sd[key + ".bias"] = torch.zeros(c)
            sd[key + ".running_mean"] = torch.zeros(c)
            sd[key + ".running_var"] = torch.ones(c)
        sd[key + ".num_batches_tracked"] = torch.tensor(0)

    for sl, pairs in _VGG.items():
        for ci, bi in pairs:
            cin, cout = _VGG_CH[ci]
            conv(f"basenet.{sl}.{ci}", cin, cout, 3)
            bn(f"basenet.{sl}.{bi}", cout)
    conv("basenet.slice5.1", 512, 1024, 3)
    conv("basenet.slice5.2", 1024, 1024, 1)
    for name, (i, m, o) in _UP.items():
        conv(f"{name}.conv.0", i + m, m, 1)
        bn(f"{name}.conv.1", m)
        conv(f"{name}.conv.3", m, o, 3)
        bn(f"{name}.conv.4", o)
    for idx, (cin, cout, k) in _CLS.items():
        conv(f"conv_cls.{idx}", cin, cout, k)
    return sd


def glyph_craft_state(seed=0, text_gain=3.0, link_gain=1.2, bf16_round=True):
    """Synthetic CRAFT weights that emit text-like score maps.  Random-init CRAFT outputs lie within +-0.03 and yield
    no boxes at any preset threshold (SURVEY.md hard part 6), so on top of the random initialisation one channel is
    wired, layer by layer, to carry the page's 'ink' to the heads: conv1_1 ch0 = relu(s*(b - g - r - 1)) — 1 on the
    synthetic pages' dark-blue ink (oracle/synth.py INK_BGR), 0 on white paper AND on the black zero-padding of the
    canvas (imgproc.py:60-63 pads before normalisation, so the pad is -1.0 = 'black') — then identity
    centre taps through conv1_2 / conv2_1 / conv2_2 (the 2x2 max-pool dilates thin strokes), the U-Net skip of
    upconv4, four 3x3 box blurs (upconv4.conv.3, conv_cls.0/2/4) and the last 1x1 with gains `text_gain` /
    `link_gain` — the network-computed analogue of the 'injected oracle maps' text = 3*blur(1-gray),
    link = 1.2*blur of SURVEY.md §8d.  Every other weight stays random, so all layers do real work; a letter page
    gives one component per word (~500).  The same state dict goes to the oracle and the device."""
    sd = synth_craft_state(seed, random_bn=False, bf16_round=bf16_round)

    def rnd(t):
        return t.to(torch.bfloat16).float() if bf16_round else t

    def wire(key, out_ch, in_ch, kernel):
        w = sd[key + ".weight"]
        w[out_ch] = 0
        k = torch.as_tensor(kernel, dtype=torch.float32)
        w[out_ch, in_ch] = rnd(k.reshape(w.shape[2], w.shape[3]))
        sd[key + ".bias"][out_ch] = 0

    c1 = [[0, 0, 0], [0, 1, 0], [0, 0, 0]]
    box = [[1 / 9.0] * 3] * 3
    w = sd["basenet.slice1.0.weight"]
    w[0] = 0
    s_ink = 1.0 / ((160 - 127.5) / 127.5 + 1.0)            # full ink (160, 0, 0) -> 1.0
    w[0, :, 1, 1] = rnd(torch.tensor([s_ink, -s_ink, -s_ink]))
    sd["basenet.slice1.0.bias"][0] = -s_ink
    for key in ("basenet.slice1.3", "basenet.slice1.7", "basenet.slice1.10"):
        wire(key, 0, 0, c1)
    wire("upconv4.conv.0", 0, 64, [[1.0]])             # input = cat([upsampled y (64 ch), relu2_2 (128 ch)])
    wire("upconv4.conv.3", 0, 0, box)
    for key in ("conv_cls.0", "conv_cls.2", "conv_cls.4"):
        wire(key, 0, 0, box)
    wire("conv_cls.6", 0, 0, [[1.0]])
    sd["conv_cls.8.weight"][0, 0, 0, 0] = float(rnd(torch.tensor(text_gain)))
    sd["conv_cls.8.weight"][1, 0, 0, 0] = float(rnd(torch.tensor(link_gain)))
    return sd


# ------------------------------------------------------------------------------------------------- TrOCR
BOS, PAD, EOS, UNK = 0, 1, 2, 3


@dataclass
class TrocrConfig:
    enc_dim: int = 768
    enc_layers: int = 12
    enc_heads: int = 12
    enc_ffn: int = 3072
    dec_dim: int = 1024
    dec_layers: int = 12
    dec_heads: int = 16
    dec_ffn: int = 4096
    vocab: int = 50265
    img: int = 384
    patch: int = 16
    max_positions: int = 1024

    @property
    def tokens(self):
        return (self.img // self.patch) ** 2 + 1


def trocr_base():
    return TrocrConfig()


def trocr_large():
    return TrocrConfig(enc_dim=1024, enc_layers=24, enc_heads=16, enc_ffn=4096)


def trocr_tiny(vocab=1000):
    """Same structure (head dim 64, 577 tokens) at toy widths — for fast CPU/GPU parity tests."""
    return TrocrConfig(enc_dim=128, enc_layers=2, enc_heads=2, enc_ffn=256, dec_dim=128, dec_layers=2, dec_heads=2,
                       dec_ffn=256, vocab=vocab)


# ------------------------------------------------------------------------------------------------- weights
def synth_trocr_state(cfg, seed=0, round_to=torch.float16, out_scale=0.25):
    """Random-init weights following timm's / fairseq's initialisers (trunc-normal 0.02 ViT linears, xavier-uniform
    decoder linears, N(0, d^-0.5) embeddings), rounded once to `round_to` so the oracle and the device share the
    exact same values.  The decoder's residual-writing projections (out_proj, fc2) are scaled by `out_scale` so the
    residual stream keeps the position signal that calibrate_eos() relies on."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def rnd(t):
        return t.to(round_to).float() if round_to is not None else t

    def tn(*shape, std=0.02):
        return rnd(torch.empty(*shape).normal_(0, std, generator=g).clamp_(-2 * std, 2 * std))

    def xavier(out_f, in_f, gain=1.0):
        b = gain * math.sqrt(6.0 / (in_f + out_f))
        return rnd(torch.empty(out_f, in_f).uniform_(-b, b, generator=g))

    def ln(key, d):
        sd[key + ".weight"] = rnd(1.0 + 0.1 * torch.empty(d).normal_(0, 1, generator=g))
        sd[key + ".bias"] = rnd(0.05 * torch.empty(d).normal_(0, 1, generator=g))

    def bias(d, std=0.02):
        return rnd(torch.empty(d).normal_(0, std, generator=g))

    D = cfg.enc_dim
    e = "encoder.deit."
    sd[e + "patch_embed.proj.weight"] = tn(D, 3, cfg.patch, cfg.patch)
    sd[e + "patch_embed.proj.bias"] = bias(D)
    sd[e + "cls_token"] = tn(1, 1, D)
    sd[e + "pos_embed"] = tn(1, cfg.tokens, D)
    for i in range(cfg.enc_layers):
        b = f"{e}blocks.{i}."
        ln(b + "norm1", D)
        sd[b + "attn.qkv.weight"] = tn(3 * D, D, std=0.05)
        sd[b + "attn.proj.weight"] = tn(D, D)
        sd[b + "attn.proj.bias"] = bias(D)
        ln(b + "norm2", D)
        sd[b + "mlp.fc1.weight"] = tn(cfg.enc_ffn, D)
        sd[b + "mlp.fc1.bias"] = bias(cfg.enc_ffn)
        sd[b + "mlp.fc2.weight"] = tn(D, cfg.enc_ffn)
        sd[b + "mlp.fc2.bias"] = bias(D)
    ln(e + "norm", D)

    H = cfg.dec_dim
    emb = torch.empty(cfg.vocab, H).normal_(0, H ** -0.5, generator=g)
    emb[PAD] = 0
    sd["decoder.embed_tokens.weight"] = rnd(emb)
    for i in range(cfg.dec_layers):
        b = f"decoder.layers.{i}."
        for name, kdim in (("self_attn", H), ("encoder_attn", D)):
            sd[b + name + ".q_proj.weight"] = xavier(H, H, 2 ** -0.5)
            sd[b + name + ".k_proj.weight"] = xavier(H, kdim, 2 ** -0.5)
            sd[b + name + ".v_proj.weight"] = xavier(H, kdim, 2 ** -0.5)
            sd[b + name + ".out_proj.weight"] = xavier(H, H, out_scale)
            for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
                sd[b + name + f".{p}.bias"] = bias(H)
        ln(b + "self_attn_layer_norm", H)
        ln(b + "encoder_attn_layer_norm", H)
        sd[b + "fc1.weight"] = xavier(cfg.dec_ffn, H)
        sd[b + "fc1.bias"] = bias(cfg.dec_ffn)
        sd[b + "fc2.weight"] = xavier(H, cfg.dec_ffn, out_scale)
        sd[b + "fc2.bias"] = bias(H)
        ln(b + "final_layer_norm", H)
    wout = torch.empty(cfg.vocab, H).normal_(0, H ** -0.5, generator=g)
    sd["decoder.output_projection.weight"] = rnd(wout)
    return sd


def apply_eos_row(sd, name="trocr_base_seed0", round_to=torch.float16):
    """Installs the pre-computed EOS row of the output projection (synthetic/eos_row_<name>.npy, produced by
    tools/make_eos_rows.py with the oracle's calibrate_eos on real encoder states of synthetic word crops)."""
    row = torch.from_numpy(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"eos_row_{name}.npy")))
    w = sd["decoder.output_projection.weight"]
    assert row.shape[0] == w.shape[1]
    w[EOS] = row.to(round_to).float() if round_to is not None else row
    return sd


# RefineNet (marie/models/craft/refinenet.py:15-55): Sequential indices of the convolutions / batch norms
_REFINE = {"last_conv": [(0, 1, 34, 64, 3), (3, 4, 64, 64, 3), (6, 7, 64, 64, 3)]}
for _k in range(1, 5):
    _REFINE[f"aspp{_k}"] = [(0, 1, 64, 128, 3), (3, 4, 128, 128, 1), (6, None, 128, 1, 1)]


def synth_refine_state(seed=0, random_bn=False, round_to=torch.float16, out_gain=1.0, out_bias=0.0):
    """RefineNet state dict as its own init_weights leaves it (xavier-uniform convolutions, zero biases, BN gamma 1 /
    beta 0 / running stats 0 / 1; vgg16_bn.py:10-21), convolution weights rounded once to `round_to`.  `out_gain` /
    `out_bias` rescale the four final 1x1 convolutions (a random-init refiner emits values near zero, below every link
    threshold)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, layers in _REFINE.items():
        for ci, bi, cin, cout, k in layers:
            w = torch.empty(cout, cin, k, k)
            bound = (6.0 / (cin * k * k + cout * k * k)) ** 0.5
            w.uniform_(-bound, bound, generator=g)
            last = bi is None
            if last:
                w = w * out_gain
            if round_to is not None:
                w = w.to(round_to).float()
            sd[f"{name}.{ci}.weight"] = w
            sd[f"{name}.{ci}.bias"] = torch.full((cout,), out_bias / 4.0) if last else torch.zeros(cout)
            if bi is not None:
                if random_bn:
                    sd[f"{name}.{bi}.weight"] = torch.empty(cout).uniform_(0.6, 1.4, generator=g)
                    sd[f"{name}.{bi}.bias"] = torch.empty(cout).uniform_(-0.2, 0.2, generator=g)
                    sd[f"{name}.{bi}.running_mean"] = torch.empty(cout).uniform_(-0.2, 0.2, generator=g)
                    sd[f"{name}.{bi}.running_var"] = torch.empty(cout).uniform_(0.5, 1.5, generator=g)
                else:
                    sd[f"{name}.{bi}.weight"] = torch.ones(cout)
                    sd[f"{name}.{bi}.bias"] = torch.zeros(cout)
                    sd[f"{name}.{bi}.running_mean"] = torch.zeros(cout)
                    sd[f"{name}.{bi}.running_var"] = torch.ones(cout)
                sd[f"{name}.{bi}.num_batches_tracked"] = torch.tensor(0)
    return sd
