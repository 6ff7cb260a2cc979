"""CPU restatement of the CRAFT network and its synthetic weights (TEST INFRASTRUCTURE — see oracle/__init__.py).

  craft_forward      CRAFT.forward (marie/models/craft/craft.py:59-81) over vgg16_bn.forward
                     (marie/models/craft/basenet/vgg16_bn.py:61-73) written with torch.nn.functional on a plain
                     state dict whose keys are the reference's.  Pinned against the reference module itself in
                     tests/test_oracle_vs_reference.py (same state dict -> identical outputs).
  synth_craft_state  init_weights (vgg16_bn.py:10-21): xavier-uniform convolutions, zero biases, BN gamma=1 beta=0
                     running stats 0/1; optional randomised BN statistics to exercise the BN folding of the packer.
  refine_forward     RefineNet.forward (marie/models/craft/refinenet.py:57-66): cat(y, upconv4) -> three 3x3 conv+BN+ReLU
                     -> four ASPP branches (3x3 dilation 6/12/18/24, 1x1, 1x1 -> 1 channel) -> sum.  Pinned against the
                     reference module (tests/test_oracle_vs_reference.py) and tests/golden/refine_net.npz.
  calibrate_head     the "calibrated head" of SURVEY.md §8d: random-init CRAFT emits maps within ±0.03, so the
                     last 1x1 convolution is rescaled/biased until a fixed percentile of each map crosses the
                     PSM thresholds.  Same weights go to the oracle and the device.
"""
import numpy as np
import torch
import torch.nn.functional as F

from synthetic.weights import glyph_craft_state, synth_craft_state, synth_refine_state  # noqa: F401  (generators live outside oracle/)


def _cbr(sd, x, ck, bk, relu=True, padding=1, dilation=1):
    x = F.conv2d(x, sd[ck + ".weight"], sd[ck + ".bias"], padding=padding, dilation=dilation)
    if bk is not None:
        x = F.batch_norm(x, sd[bk + ".running_mean"], sd[bk + ".running_var"], sd[bk + ".weight"], sd[bk + ".bias"],
                         False, 0.0, 1e-5)
    return F.relu(x) if relu else x


def craft_forward(sd, x):
    """x: [B,3,H,W] float32 -> (y [B,H/2,W/2,2], feature [B,32,H/2,W/2])."""
    b = "basenet."
    h = _cbr(sd, x, b + "slice1.0", b + "slice1.1")
    h = _cbr(sd, h, b + "slice1.3", b + "slice1.4")
    h = F.max_pool2d(h, 2, 2)
    h = _cbr(sd, h, b + "slice1.7", b + "slice1.8")
    relu2_2 = _cbr(sd, h, b + "slice1.10", b + "slice1.11")        # post-ReLU: in-place ReLU of slice2[12]
    h = F.max_pool2d(relu2_2, 2, 2)
    h = _cbr(sd, h, b + "slice2.14", b + "slice2.15")
    relu3_2 = _cbr(sd, h, b + "slice2.17", b + "slice2.18")
    h = _cbr(sd, relu3_2, b + "slice3.20", b + "slice3.21")
    h = F.max_pool2d(h, 2, 2)
    h = _cbr(sd, h, b + "slice3.24", b + "slice3.25")
    relu4_3 = _cbr(sd, h, b + "slice3.27", b + "slice3.28")
    h = _cbr(sd, relu4_3, b + "slice4.30", b + "slice4.31")
    h = F.max_pool2d(h, 2, 2)
    h = _cbr(sd, h, b + "slice4.34", b + "slice4.35")
    relu5_3 = _cbr(sd, h, b + "slice4.37", b + "slice4.38", relu=False)   # BN(conv5_2), no ReLU
    h = F.max_pool2d(relu5_3, 3, 1, 1)
    h = F.conv2d(h, sd[b + "slice5.1.weight"], sd[b + "slice5.1.bias"], padding=6, dilation=6)
    fc7 = F.conv2d(h, sd[b + "slice5.2.weight"], sd[b + "slice5.2.bias"])

    def double_conv(name, y):
        y = _cbr(sd, y, f"{name}.conv.0", f"{name}.conv.1", padding=0)
        return _cbr(sd, y, f"{name}.conv.3", f"{name}.conv.4")

    y = double_conv("upconv1", torch.cat([fc7, relu5_3], 1))
    y = F.interpolate(y, size=relu4_3.shape[2:], mode="bilinear", align_corners=False)
    y = double_conv("upconv2", torch.cat([y, relu4_3], 1))
    y = F.interpolate(y, size=relu3_2.shape[2:], mode="bilinear", align_corners=False)
    y = double_conv("upconv3", torch.cat([y, relu3_2], 1))
    y = F.interpolate(y, size=relu2_2.shape[2:], mode="bilinear", align_corners=False)
    feature = double_conv("upconv4", torch.cat([y, relu2_2], 1))
    y = F.relu(F.conv2d(feature, sd["conv_cls.0.weight"], sd["conv_cls.0.bias"], padding=1))
    y = F.relu(F.conv2d(y, sd["conv_cls.2.weight"], sd["conv_cls.2.bias"], padding=1))
    y = F.relu(F.conv2d(y, sd["conv_cls.4.weight"], sd["conv_cls.4.bias"], padding=1))
    y = F.relu(F.conv2d(y, sd["conv_cls.6.weight"], sd["conv_cls.6.bias"]))
    y = F.conv2d(y, sd["conv_cls.8.weight"], sd["conv_cls.8.bias"])
    return y.permute(0, 2, 3, 1), feature


def calibrate_head(sd, y_uncal, low_text=0.3, link_threshold=0.45, text_pct=98.5, link_pct=97.0, peak=1.0):
    """Rescales conv_cls.8 in place so that, on the calibration page, the `pct`-th percentile of each map equals
    its threshold and the 99.9th percentile reaches ~`peak` (so max(text) >= text_threshold for real blobs).
    y_uncal: [H,W,2] output of craft_forward with the un-calibrated head.  Returns the (scale, bias) pairs."""
    out = []
    for ch, (thr, pct) in enumerate([(low_text, text_pct), (link_threshold, link_pct)]):
        v = y_uncal[..., ch].reshape(-1).double().numpy()
        p_lo, p_hi = np.percentile(v, pct), np.percentile(v, 99.9)
        scale = (peak - thr) / max(p_hi - p_lo, 1e-12)
        bias = thr - scale * p_lo
        w = sd["conv_cls.8.weight"][ch] * scale
        sd["conv_cls.8.weight"][ch] = w.to(torch.bfloat16).float()
        sd["conv_cls.8.bias"][ch] = float(sd["conv_cls.8.bias"][ch] * scale + bias)
        out.append((float(scale), float(bias)))
    return out


def refine_forward(sd, y, feature):
    """y: [B,H,W,2] (CRAFT.forward output), feature: [B,32,H,W] -> refined link map [B,H,W,1]."""
    h = torch.cat([y.permute(0, 3, 1, 2), feature], 1)
    h = _cbr(sd, h, "last_conv.0", "last_conv.1")
    h = _cbr(sd, h, "last_conv.3", "last_conv.4")
    h = _cbr(sd, h, "last_conv.6", "last_conv.7")
    out = None
    for k, d in ((1, 6), (2, 12), (3, 18), (4, 24)):
        a = _cbr(sd, h, f"aspp{k}.0", f"aspp{k}.1", padding=d, dilation=d)
        a = _cbr(sd, a, f"aspp{k}.3", f"aspp{k}.4", padding=0)
        a = F.conv2d(a, sd[f"aspp{k}.6.weight"], sd[f"aspp{k}.6.bias"])
        out = a if out is None else out + a
    return out.permute(0, 2, 3, 1)
