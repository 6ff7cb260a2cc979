"""CPU restatement of the TrOCR recogniser in plain PyTorch fp32 (TEST INFRASTRUCTURE — see oracle/__init__.py).

Pinning status, part by part (tests/test_oracle_vs_reference.py, run against /root/reference in the build container):
  encoder   PINNED on reference-held code: `encoder_forward` == the reference's own
            AdaptedVisionTransformer.forward_features (marie/models/unilm/trocr/deit.py:105-146) running on the ViT blocks
            the reference vendors (marie/boxes/dit/ditod/deit.py:44-167; only `timm.models.layers` helpers are stubbed),
            strict state-dict load, tiny widths and one crop at the beit_base_patch16_384 factory geometry (deit.py:323-329).
  search    PINNED on the reference's own loop: `generate` == TextRecognitionGenerator._generate (generator.py:11-374)
            executed from the reference file on the same incremental decoder — tokens, scores and positional scores of every
            finalised hypothesis, beams 1 / 2 / 3 / 5, batch compaction and max_len EOS forcing.  fairseq's BeamSearch.step /
            finalize_hypos, which that loop calls, are restated (oracle/ref_loader.load_generator): fairseq is absent.
  text      PINNED: GPT-2 BPE decode == the reference's GPT2BPEEnhancedSpace.decode (bpe.py:59-67).
  decoder   PARITY UNPINNED against the reference itself: the layer arithmetic is fairseq's TransformerDecoder (unpinned git
            HEAD in the reference's requirements, not vendored, not installed; no golden vectors).  Pinned by the reference's
            construction site and arch defaults instead — trocr_models.py:142-147, :423-447 (dim 1024, ffn 4096, 16 heads,
            12 layers, post-LN, ReLU, sinusoidal positions, embed scale sqrt(d), untied bias-free output projection,
            cross-attention kdim = encoder dim), TrOCREncoder.forward (trocr_models.py:508-524) — and by an independent
            cross-check against HuggingFace transformers' port on identical weights (tests/test_oracle_trocr.py: ViTModel /
            TrOCRForCausalLM agree to 1e-4).

State-dict keys are fairseq's (`encoder.deit.*`, `decoder.*`), so a real TrOCR checkpoint's `model` dict can be
passed to the same functions and to marie-icr_b200/weights.py:pack_trocr.
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from synthetic.weights import (BOS, EOS, PAD, UNK, TrocrConfig, synth_trocr_state, trocr_base, trocr_large,  # noqa: F401
                               trocr_tiny)


def calibrate_eos(sd, cfg, eos_step=6, samples=6, seed=0, margin=2.0, round_to=torch.float16, enc=None, horizon=32):
    """A random-init decoder never emits EOS (1 chance in V per step), so every hypothesis would run to max_len=200.
    Like the CRAFT 'calibrated head'  the EOS row of the output projection is fitted, in place, so that
    its logit ramps through the level of the best competing logit at step eos_step (ridge regression on a few
    greedy roll-outs over the encoder states `enc` — pass real encoder outputs of a few sample crops; random
    states are only a fallback), scaled so EOS wins the arg-max around step `eos_step`.  The roll-out runs `horizon` steps past eos_step so the
    direction is dominated by the slowly varying position components and EOS keeps winning at every later step
    (no straggler hypotheses that run to max_len).  The same
    weights go to the oracle and the device, so parity is unaffected.  Returns the scale used."""
    g = torch.Generator().manual_seed(seed + 77)
    if enc is None:
        enc = torch.randn(samples, cfg.tokens, cfg.enc_dim, generator=g)
    samples = enc.shape[0]
    W = sd["decoder.output_projection.weight"]
    W[EOS] = 0
    st = DecoderState(sd, cfg, enc)
    tok = torch.full((samples,), EOS, dtype=torch.long)
    hs = []
    with torch.no_grad():
        for t in range(eos_step + horizon):
            logits, h = st.step(tok, t, return_hidden=True)
            hs.append(h)
            logits[:, PAD] = -math.inf
            logits[:, EOS] = -math.inf
            tok = logits.argmax(-1)
        hs = torch.stack(hs)                                   # [T, S, H]
        T, S, H = hs.shape
        other = (hs @ W.t()).max(-1).values                    # [T, S] best competing logit
        # ridge regression of an EOS logit that ramps through the competitors' level at step eos_step:
        # target = other + clip(slope * (t - eos_step + 0.5), -6, 6)
        ramp = (margin * (torch.arange(T, dtype=torch.float) - eos_step + 0.5)).clamp(-6, 6)
        y = (other + ramp[:, None]).reshape(-1).double()
        Hm = hs.reshape(T * S, H).double()
        G = Hm @ Hm.t()
        lam = 1e-2 * float(G.diagonal().mean())
        w = Hm.t() @ torch.linalg.solve(G + lam * torch.eye(T * S, dtype=torch.double), y)
        W[EOS] = w.float().to(round_to).float() if round_to is not None else w.float()
        alpha = float(w.norm())
    return alpha


# ------------------------------------------------------------------------------------------------- encoder
def _ln(x, sd, key, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def encoder_forward(sd, cfg, imgs):
    """imgs [B,3,384,384] f32 -> [B, 577, D]  (deit.py:105-146 with timm 0.6.12 Block/Attention/Mlp semantics)."""
    e = "encoder.deit."
    x = F.conv2d(imgs, sd[e + "patch_embed.proj.weight"], sd[e + "patch_embed.proj.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)                                   # B, 576, D (row-major patches)
    x = torch.cat([sd[e + "cls_token"].expand(x.shape[0], -1, -1), x], 1) + sd[e + "pos_embed"]
    B, T, D = x.shape
    h, dh = cfg.enc_heads, D // cfg.enc_heads
    for i in range(cfg.enc_layers):
        b = f"{e}blocks.{i}."
        y = _ln(x, sd, b + "norm1", 1e-6)
        qkv = F.linear(y, sd[b + "attn.qkv.weight"]).reshape(B, T, 3, h, dh).permute(2, 0, 3, 1, 4)
        att = (qkv[0] @ qkv[1].transpose(-2, -1)) * dh ** -0.5
        y = (att.softmax(-1) @ qkv[2]).transpose(1, 2).reshape(B, T, D)
        x = x + F.linear(y, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        y = _ln(x, sd, b + "norm2", 1e-6)
        y = F.gelu(F.linear(y, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))
        x = x + F.linear(y, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
    return _ln(x, sd, e + "norm", 1e-6)


# ------------------------------------------------------------------------------------------------- decoder
def sinusoidal_table(n, dim, padding_idx=PAD):
    """fairseq SinusoidalPositionalEmbedding.get_embedding: [sin | cos] halves, zero row at padding_idx."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    ang = torch.arange(n, dtype=torch.float)[:, None] * freq[None]
    t = torch.cat([ang.sin(), ang.cos()], 1)
    t[padding_idx] = 0
    return t


class DecoderState:
    """Incremental state: per layer self-attention K/V [rows, heads, t, dh] and static cross-attention K/V."""

    def __init__(self, sd, cfg, enc_out, beam=1):
        """enc_out [B, T, E]; rows of the state are B*beam (row = sentence*beam + beam slot).  The static cross-attention
        K/V are computed and kept once per SENTENCE: fairseq computes them on the beam-expanded encoder output
        (generator.py:68-70) and reorders them every step, but all beams of a sentence share identical encoder states
        and a reorder never crosses sentences, so the values are the same."""
        self.sd, self.cfg, self.beam = sd, cfg, beam
        H, h = cfg.dec_dim, cfg.dec_heads
        self.dh = H // h
        B, T, _ = enc_out.shape
        self.cross = []
        for i in range(cfg.dec_layers):
            b = f"decoder.layers.{i}.encoder_attn."
            k = F.linear(enc_out, sd[b + "k_proj.weight"], sd[b + "k_proj.bias"]).view(B, T, h, self.dh).transpose(1, 2)
            v = F.linear(enc_out, sd[b + "v_proj.weight"], sd[b + "v_proj.bias"]).view(B, T, h, self.dh).transpose(1, 2)
            self.cross.append((k, v))
        self.self_kv = [None] * cfg.dec_layers
        self.pe = sinusoidal_table(cfg.max_positions + 2, H)

    def reorder(self, idx):
        self.self_kv = [(k[idx], v[idx]) for k, v in self.self_kv]

    def step(self, tokens_last, t, return_hidden=False):
        """tokens_last [rows] (token at position t, 0-based) -> logits [rows, V] (fp32)."""
        sd, cfg = self.sd, self.cfg
        H, h, dh = cfg.dec_dim, cfg.dec_heads, self.dh
        x = math.sqrt(H) * sd["decoder.embed_tokens.weight"][tokens_last] + self.pe[PAD + 1 + t]
        R = x.shape[0]
        for i in range(cfg.dec_layers):
            b = f"decoder.layers.{i}."
            q = F.linear(x, sd[b + "self_attn.q_proj.weight"], sd[b + "self_attn.q_proj.bias"]) * dh ** -0.5
            k = F.linear(x, sd[b + "self_attn.k_proj.weight"], sd[b + "self_attn.k_proj.bias"]).view(R, h, 1, dh)
            v = F.linear(x, sd[b + "self_attn.v_proj.weight"], sd[b + "self_attn.v_proj.bias"]).view(R, h, 1, dh)
            if self.self_kv[i] is not None:
                k = torch.cat([self.self_kv[i][0], k], 2)
                v = torch.cat([self.self_kv[i][1], v], 2)
            self.self_kv[i] = (k, v)
            a = (q.view(R, h, 1, dh) @ k.transpose(-2, -1)).softmax(-1) @ v
            x = _ln(x + F.linear(a.reshape(R, H), sd[b + "self_attn.out_proj.weight"], sd[b + "self_attn.out_proj.bias"]),
                    sd, b + "self_attn_layer_norm", 1e-5)
            q = F.linear(x, sd[b + "encoder_attn.q_proj.weight"], sd[b + "encoder_attn.q_proj.bias"]) * dh ** -0.5
            ck, cv = self.cross[i]                                     # [B, h, T, dh], shared by the beams of a sentence
            q4 = q.view(R // self.beam, self.beam, h, dh).transpose(1, 2)
            a = ((q4 @ ck.transpose(-2, -1)).softmax(-1) @ cv).transpose(1, 2)
            x = _ln(x + F.linear(a.reshape(R, H), sd[b + "encoder_attn.out_proj.weight"],
                                 sd[b + "encoder_attn.out_proj.bias"]), sd, b + "encoder_attn_layer_norm", 1e-5)
            y = F.linear(F.relu(F.linear(x, sd[b + "fc1.weight"], sd[b + "fc1.bias"])), sd[b + "fc2.weight"], sd[b + "fc2.bias"])
            x = _ln(x + y, sd, b + "final_layer_norm", 1e-5)
        logits = F.linear(x, sd["decoder.output_projection.weight"])
        return (logits, x) if return_hidden else logits


# ------------------------------------------------------------------------------------------------- search
def generate(sd, cfg, enc_out, beam=1, max_len_b=200, min_len=1, forced=None, trace=None, margins=None):
    """fairseq sequence generation as configured by the reference (generator.py:11-374; task.py:165-276):
    returns, per input row, the list of finalised hypotheses sorted by score (desc), each a dict
    {tokens: LongTensor (ends with EOS), score: float (sum log-prob / length), positional_scores}.

    Sentences are independent in fairseq's search, so finished sentences are simply frozen here instead of being
    removed from the batch (generator.py:262-297 only compacts the tensors).
    `forced`: optional [bsz, L] token matrix — teacher forcing for parity checks: the search is replaced by taking
    forced[:, step] (beam must be 1) while `trace` collects (step, lprobs) for margin analysis.
    `margins`: optional list that receives, per input row, the smallest gap between neighbouring candidate scores the
    search ever had to order while the sentence was live (beam 1: top-1 vs top-2 log-prob; beam b: the top 2b+1
    cumulative scores) — the parity tests' margin protocol: a 16-bit implementation may legitimately order a closer
    call differently."""
    bsz = enc_out.shape[0]
    V = cfg.vocab
    max_len = min(int(max_len_b), cfg.max_positions - 1)
    rows = bsz * beam
    st = DecoderState(sd, cfg, enc_out, beam)
    tokens = torch.full((rows, max_len + 2), PAD, dtype=torch.long)
    tokens[:, 0] = EOS
    scores = torch.zeros(rows, max_len + 1)
    ignore = torch.zeros(bsz, beam, dtype=torch.bool)
    finalized = [[] for _ in range(bsz)]
    finished = [False] * bsz
    cand = 2 * beam
    for step in range(max_len + 1):
        logits = st.step(tokens[:, step], step)
        lprobs = F.log_softmax(logits.float(), -1)
        lprobs[lprobs != lprobs] = -math.inf
        lprobs[:, PAD] = -math.inf
        if step >= max_len:
            lprobs[:, :EOS] = -math.inf
            lprobs[:, EOS + 1:] = -math.inf
        elif step < min_len:
            lprobs[:, EOS] = -math.inf
        if trace is not None:
            trace.append((step, lprobs.clone()))
        lp = lprobs.view(bsz, beam, V)
        if forced is not None:
            assert beam == 1
            if step >= forced.shape[1]:
                break
            tok = forced[:, step]
            tokens[:, step + 1] = tok
            scores[:, step] = (scores[:, step - 1] if step else 0) + lp[torch.arange(bsz), 0, tok]
            continue
        if step == 0:
            lp = lp[:, :1, :]
        else:
            lp = lp + scores.view(bsz, beam, -1)[:, :, step - 1].unsqueeze(-1)
        flat = lp.reshape(bsz, -1)
        c_scores, c_idx = torch.topk(flat, k=min(cand, flat.shape[1] - 1))
        if margins is not None:
            if not margins:
                margins.extend([math.inf] * bsz)
            top = torch.topk(flat, k=min((1 if beam == 1 else cand) + 1, flat.shape[1])).values
            gap = (top[:, :-1] - top[:, 1:]).min(-1).values
            for s in range(bsz):
                if not finished[s] and math.isfinite(float(gap[s])):
                    margins[s] = min(margins[s], float(gap[s]))
        c_beam, c_tok = c_idx // V, c_idx % V
        new_tokens, new_scores, reorder = tokens.clone(), scores.clone(), torch.arange(rows)
        for s in range(bsz):
            if finished[s]:
                continue
            eos_mask = (c_tok[s] == EOS) & (c_scores[s] != -math.inf)
            eos_mask[:beam] &= ~ignore[s]
            for c in range(beam):                          # only EOS among the top `beam` candidates finalises
                if eos_mask[c] and len(finalized[s]) < beam:
                    src = s * beam + int(c_beam[s, c])
                    toks = torch.cat([tokens[src, 1:step + 1], torch.tensor([EOS])])
                    pos = torch.cat([scores[src, :step], c_scores[s, c:c + 1]])
                    pos[1:] = pos[1:] - pos[:-1].clone()
                    finalized[s].append(dict(tokens=toks, score=float(c_scores[s, c]) / (step + 1),
                                             positional_scores=pos))
            if len(finalized[s]) == beam or step == max_len:
                finished[s] = True
                continue
            eos_mask[:beam] |= ignore[s]
            active = eos_mask.long() * cand + torch.arange(len(eos_mask))
            vals, hyp = torch.topk(active, k=beam, largest=False)
            ignore[s] = vals >= cand
            for j in range(beam):
                c = int(hyp[j])
                src = s * beam + int(c_beam[s, c])
                dst = s * beam + j
                new_tokens[dst, :step + 1] = tokens[src, :step + 1]
                new_tokens[dst, step + 1] = c_tok[s, c]
                new_scores[dst, :step] = scores[src, :step]
                new_scores[dst, step] = c_scores[s, c]
                reorder[dst] = src
        tokens, scores = new_tokens, new_scores
        if all(finished):
            break
        st.reorder(reorder)
    if forced is not None:
        return tokens, scores
    for s in range(bsz):
        order = sorted(range(len(finalized[s])), key=lambda i: -finalized[s][i]["score"])
        finalized[s] = [finalized[s][i] for i in order]
    return finalized


def recognize(sd, cfg, imgs, beam=1, max_len_b=200):
    """imgs [B,3,384,384] f32 -> list of (token ids without EOS, confidence = exp(score)) like get_text
    (trocr_ocr_processor.py:142-180) before detokenisation."""
    enc = encoder_forward(sd, cfg, imgs)
    out = []
    for hyps in generate(sd, cfg, enc, beam=beam, max_len_b=max_len_b):
        top = hyps[0]
        out.append((top["tokens"][:-1].tolist(), math.exp(top["score"])))
    return out
