"""TrOcrProcessorB200 — drop-in for TrOcrProcessor (marie/document/trocr_ocr_processor.py:188-367): same constructor
arguments and errors, `recognize_from_fragments(images) -> [{"confidence", "id", "text"}]` in input order, `recognize`
inherited from the OcrProcessor mirror.  All fragments of a call travel to the device in one copy, are resized /
normalised / packed by K9 and decoded by the device-resident search (no per-crop H2D, no per-hypothesis D2H).
"""
import math
import os

import numpy as np
import torch

from . import weights as _weights
from .bpe import SyntheticDetokenizer
from .pipeline import PagePipeline
from .plugin_api import OcrProcessor


class TrOcrProcessorB200(OcrProcessor):
    def __init__(self, work_dir="/tmp/icr", model_name_or_path=None, cuda=True, *, state_dict=None, config=None,
                 beam=3, max_len_b=200, detokenizer=None, pipeline=None, device=0, **kwargs):
        """beam defaults to 3 like the reference (:228).  state_dict/config: fairseq TrOCR weights and geometry; when
        omitted, `model_name_or_path` must point at a fairseq checkpoint (:199-222)."""
        super().__init__(work_dir, cuda, **kwargs)
        if cuda and not torch.cuda.is_available():
            raise RuntimeError("CUDA specified but no cuda devices found ")        # same error as :221-222
        if not cuda:
            raise RuntimeError("TrOcrProcessorB200 has no CPU path: cuda=True and a B200 are required")
        self.pipeline = pipeline or PagePipeline(device=device)
        if state_dict is None and not self.pipeline.has_trocr:
            if not model_name_or_path or not os.path.exists(model_name_or_path):
                raise FileNotFoundError(f"Model not found : {model_name_or_path}")   # :216-217
            ckpt = torch.load(model_name_or_path, map_location="cpu", weights_only=False)
            state_dict = ckpt["model"]
            if config is None:
                config = _config_from_state(state_dict)
        if state_dict is not None:
            self.pipeline.load_trocr(_weights.pack_trocr(state_dict, config, self.pipeline.dtype))
        self.beam, self.max_len_b = beam, max_len_b
        self.detok = detokenizer or SyntheticDetokenizer()

    def is_available(self) -> bool:
        return self.pipeline.has_trocr

    def recognize_from_fragments(self, images, **kwargs):
        """One {"confidence", "id": "img-k", "text"} per fragment, same order (:251-367); text upper-cased, confidence
        = round(round(exp(score), 6), 4) (:159-160,338-341)."""
        if len(images) == 0:
            return []
        frags = [np.asarray(f) for f in images]
        tokens, lengths, scores = self.pipeline.recognize_fragments(frags, beam=self.beam, max_len_b=self.max_len_b,
                                                                    out_ld=min(self.max_len_b + 1, 64))
        tokens, lengths, scores = tokens.cpu().numpy(), lengths.cpu().numpy(), scores.cpu().numpy()
        results = []
        for k in range(len(frags)):
            n = int(lengths[k])
            text = self.detok.decode(tokens[k, :n].tolist()).upper()
            conf = round(round(math.exp(float(scores[k])), 6), 4) if n else 0.0
            results.append({"confidence": conf, "id": f"img-{k}", "text": text})
        return results


def _config_from_state(sd):
    """Geometry from a fairseq TrOCR state dict (arch tables of marie/models/unilm/trocr/trocr_models.py:423-447)."""
    from types import SimpleNamespace
    D = sd["encoder.deit.pos_embed"].shape[-1]
    enc_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("encoder.deit.blocks."))
    dec_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers."))
    H = sd["decoder.embed_tokens.weight"].shape[1]
    return SimpleNamespace(enc_dim=D, enc_layers=enc_layers, enc_heads=D // 64,
                           enc_ffn=sd["encoder.deit.blocks.0.mlp.fc1.weight"].shape[0], dec_dim=H, dec_layers=dec_layers,
                           dec_heads=H // 64, dec_ffn=sd["decoder.layers.0.fc1.weight"].shape[0],
                           vocab=sd["decoder.embed_tokens.weight"].shape[0],
                           tokens=sd["encoder.deit.pos_embed"].shape[1], max_positions=1024)
