"""TrOCR oracle (oracle/trocr.py) cross-checked against HuggingFace transformers' port of the same architecture on
identical weights — the independent second opinion of SURVEY.md §8c (the reference's own fairseq/timm code is not
available: parity of this half is 'unpinned' against the reference itself).  Also exercises the search semantics
on hand-checkable cases."""
import math

import pytest
import torch

from oracle import trocr


@pytest.fixture(scope="module")
def tiny():
    cfg = trocr.trocr_tiny(vocab=600)
    sd = trocr.synth_trocr_state(cfg, 1, round_to=None)
    trocr.calibrate_eos(sd, cfg, round_to=None)
    return cfg, sd


def test_encoder_matches_hf_vit(tiny):
    from transformers import ViTConfig, ViTModel
    cfg, sd = tiny
    hf = ViTModel(ViTConfig(hidden_size=cfg.enc_dim, num_hidden_layers=cfg.enc_layers,
                            num_attention_heads=cfg.enc_heads, intermediate_size=cfg.enc_ffn, image_size=384,
                            patch_size=16, qkv_bias=False, layer_norm_eps=1e-6, hidden_act="gelu",
                            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0), add_pooling_layer=False).eval()
    e = "encoder.deit."
    m = {"embeddings.cls_token": sd[e + "cls_token"], "embeddings.position_embeddings": sd[e + "pos_embed"],
         "embeddings.patch_embeddings.projection.weight": sd[e + "patch_embed.proj.weight"],
         "embeddings.patch_embeddings.projection.bias": sd[e + "patch_embed.proj.bias"],
         "layernorm.weight": sd[e + "norm.weight"], "layernorm.bias": sd[e + "norm.bias"]}
    D = cfg.enc_dim
    for i in range(cfg.enc_layers):
        b, h = f"{e}blocks.{i}.", f"encoder.layer.{i}."
        qkv = sd[b + "attn.qkv.weight"]
        m[h + "attention.attention.query.weight"] = qkv[:D]
        m[h + "attention.attention.key.weight"] = qkv[D:2 * D]
        m[h + "attention.attention.value.weight"] = qkv[2 * D:]
        m[h + "attention.output.dense.weight"] = sd[b + "attn.proj.weight"]
        m[h + "attention.output.dense.bias"] = sd[b + "attn.proj.bias"]
        for a, c in (("layernorm_before", "norm1"), ("layernorm_after", "norm2"), ("intermediate.dense", "mlp.fc1"),
                     ("output.dense", "mlp.fc2")):
            m[h + a + ".weight"] = sd[b + c + ".weight"]
            m[h + a + ".bias"] = sd[b + c + ".bias"]
    missing, unexpected = hf.load_state_dict(m, strict=False)
    assert not unexpected and all("bias" in k and "attention.attention" in k for k in missing), (missing, unexpected)
    torch.manual_seed(0)
    imgs = torch.rand(2, 3, 384, 384) * 2 - 1
    with torch.no_grad():
        ref = hf(pixel_values=imgs).last_hidden_state
        out = trocr.encoder_forward(sd, cfg, imgs)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4), (out - ref).abs().max()


def test_decoder_matches_hf_trocr(tiny):
    from transformers import TrOCRConfig, TrOCRForCausalLM
    cfg, sd = tiny
    hf = TrOCRForCausalLM(TrOCRConfig(
        vocab_size=cfg.vocab, d_model=cfg.dec_dim, decoder_layers=cfg.dec_layers,
        decoder_attention_heads=cfg.dec_heads, decoder_ffn_dim=cfg.dec_ffn, activation_function="relu",
        max_position_embeddings=cfg.max_positions, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        use_learned_position_embeddings=False, scale_embedding=True, layernorm_embedding=False,
        cross_attention_hidden_size=cfg.enc_dim, tie_word_embeddings=False, pad_token_id=1, bos_token_id=0,
        eos_token_id=2, decoder_start_token_id=2)).eval()
    m = {}
    for k, v in sd.items():
        if k.startswith("decoder.layers.") or k == "decoder.embed_tokens.weight":
            m["model." + k] = v
    m["output_projection.weight"] = sd["decoder.output_projection.weight"]
    missing, unexpected = hf.load_state_dict(m, strict=False)
    assert not unexpected, unexpected
    assert all("embed_positions" in k for k in missing), missing
    torch.manual_seed(1)
    enc = torch.randn(3, cfg.tokens, cfg.enc_dim)
    toks = torch.randint(4, cfg.vocab, (3, 7))
    toks[:, 0] = trocr.EOS
    with torch.no_grad():
        ref = hf(input_ids=toks, encoder_hidden_states=enc).logits            # [3, 7, V]
        st = trocr.DecoderState(sd, cfg, enc)
        for t in range(7):
            out = st.step(toks[:, t], t)
            assert torch.allclose(out, ref[:, t], rtol=1e-4, atol=2e-4), (t, (out - ref[:, t]).abs().max())


def test_greedy_is_argmax_chain_and_forced_replay(tiny):
    cfg, sd = tiny
    torch.manual_seed(2)
    enc = torch.randn(4, cfg.tokens, cfg.enc_dim)
    with torch.no_grad():
        hyps = trocr.generate(sd, cfg, enc, beam=1, max_len_b=40)
        for s, h in enumerate(hyps):
            assert len(h) == 1 and h[0]["tokens"][-1] == trocr.EOS
            toks = h[0]["tokens"]
            trace = []
            tokens, scores = trocr.generate(sd, cfg, enc[s:s + 1], beam=1, max_len_b=40, forced=toks[None], trace=trace)
            for step, lp in trace[:len(toks)]:
                assert int(lp[0].argmax()) == int(toks[step])          # greedy = arg-max of the masked log-probs
            assert math.isclose(float(scores[0, len(toks) - 1]) / len(toks), h[0]["score"], rel_tol=1e-5)
            assert torch.allclose(h[0]["positional_scores"].sum(), scores[0, len(toks) - 1], rtol=1e-5)


def test_beam_search_properties(tiny):
    cfg, sd = tiny
    torch.manual_seed(3)
    enc = torch.randn(3, cfg.tokens, cfg.enc_dim)
    with torch.no_grad():
        g = trocr.generate(sd, cfg, enc, beam=1, max_len_b=12)
        b = trocr.generate(sd, cfg, enc, beam=4, max_len_b=12)
    for hg, hb in zip(g, b):
        assert len(hb) == 4                                   # a sentence finishes with exactly `beam` hypotheses
        sc = [h["score"] for h in hb]
        assert sc == sorted(sc, reverse=True)
        for h in hb:
            assert h["tokens"][-1] == trocr.EOS and len(h["tokens"]) <= 13 and trocr.PAD not in h["tokens"].tolist()
            assert len(h["tokens"]) >= 2                      # min_len = 1: EOS is banned at step 0
        # batch independence: a sentence decoded alone gives the same hypotheses
    with torch.no_grad():
        alone = trocr.generate(sd, cfg, enc[1:2], beam=4, max_len_b=12)[0]
    assert [h["tokens"].tolist() for h in alone] == [h["tokens"].tolist() for h in b[1]]


def test_max_len_forces_eos(tiny):
    cfg, sd = tiny
    sd2 = dict(sd)
    w = sd["decoder.output_projection.weight"].clone()
    w[trocr.EOS] = 0                                          # EOS never wins on its own
    sd2["decoder.output_projection.weight"] = w
    torch.manual_seed(4)
    enc = torch.randn(2, cfg.tokens, cfg.enc_dim)
    with torch.no_grad():
        hyps = trocr.generate(sd2, cfg, enc, beam=2, max_len_b=5)
    for h in hyps:
        assert len(h) == 2 and all(len(x["tokens"]) == 6 and x["tokens"][-1] == trocr.EOS for x in h)
