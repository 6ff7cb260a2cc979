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
from .bpe import Gpt2Detokenizer
from .checkpoint import load_fairseq_checkpoint
from .pipeline import PagePipeline
from .plugin_api import MODEL_PATH, OcrProcessor


class TrOcrProcessorB200(OcrProcessor):
    def __init__(self, work_dir="/tmp/icr", model_name_or_path=None, cuda=True, *, state_dict=None, config=None,
                 beam=3, max_len_b=200, detokenizer=None, pipeline=None, device=0, models_dir=None, **kwargs):
        """Weights come from (in this order) `state_dict` (+ `config`), a pipeline that already holds a model, or a
        fairseq checkpoint: `model_name_or_path`, else the reference's default `<model_zoo>/trocr/trocr-large-printed.pt`
        (:198-217; `models_dir` replaces <model_zoo>).  beam defaults to 3 like the reference (:228).
        detokenizer: token ids -> text.  With a checkpoint it defaults to the GPT-2 BPE the reference builds
        (task.build_bpe, :112): `gpt2_with_mask.dict.txt` / `encoder.json` next to the checkpoint or under
        `<model_zoo>/assets` (:47-49) — missing files raise instead of silently producing placeholder text.  Explicit
        weights (`state_dict=` / a loaded pipeline) need an explicit detokenizer; `SyntheticDetokenizer()` is the
        stand-in for random-init weights."""
        super().__init__(work_dir, cuda, **kwargs)
        model_zoo = models_dir or MODEL_PATH
        model_path = None
        if state_dict is None and not (pipeline is not None and pipeline.has_trocr):
            model_path = os.path.join(model_zoo, "trocr", "trocr-large-printed.pt")
            if model_name_or_path:
                assert os.path.exists(model_name_or_path)                          # :210
                model_path = model_name_or_path
            if not os.path.exists(model_path):
                raise FileNotFoundError(f"File not found : {model_path}")           # :216-217 (checked before the device, as there)
        if cuda and not torch.cuda.is_available():
            raise RuntimeError("CUDA specified but no cuda devices found ")        # same error as :221-222
        if not cuda:
            raise RuntimeError("TrOcrProcessorB200 has no CPU path: cuda=True and a B200 are required")
        self.pipeline = pipeline or PagePipeline(device=device)
        info = None
        if model_path is not None:
            state_dict, info = load_fairseq_checkpoint(model_path)
            if config is None:
                config = _config_from_state(state_dict)
        if state_dict is not None:
            self.pipeline.load_trocr(_weights.pack_trocr(state_dict, config, self.pipeline.dtype, info=info))
        self.beam, self.max_len_b = beam, max_len_b
        if detokenizer is None:
            if model_path is None:
                raise ValueError("TrOcrProcessorB200: explicit weights need an explicit detokenizer (Gpt2Detokenizer for a real "
                                 "checkpoint's state dict, SyntheticDetokenizer() for random-init test weights)")
            detokenizer = Gpt2Detokenizer.locate(os.path.dirname(os.path.abspath(model_path)), os.path.join(model_zoo, "assets"))
        self.detok = detokenizer

    def is_available(self) -> bool:
        return self.pipeline.has_trocr

    def recognize_from_fragments(self, images, **kwargs):
        """One {"confidence", "id": "img-k", "text"} per fragment, same order (:251-367); text upper-cased, confidence
        = round(round(exp(score), 6), 4) (:159-160,338-341).  Fragments are BGR ndarrays or image paths
        (marie/models/icr/memory_dataset.py:17-53)."""
        if len(images) == 0:
            return []
        frags = [_load_fragment(f) for f in images]
        out_ld = self.max_len_b + 1                      # every token of the longest possible hypothesis (+ EOS)
        tokens, lengths, scores = self.pipeline.recognize_fragments(frags, beam=self.beam, max_len_b=self.max_len_b, out_ld=out_ld)
        tokens, lengths, scores = tokens.cpu().numpy(), lengths.cpu().numpy(), scores.cpu().numpy()
        results = []
        for k in range(len(frags)):
            n = int(lengths[k])
            text = self.detok.decode(tokens[k, :n].tolist()).upper()
            conf = round(round(math.exp(float(scores[k])), 6), 4) if n else 0.0
            results.append({"confidence": conf, "id": f"img-{k}", "text": text})
        return results


def _load_fragment(f):
    """memory_dataset.py:17-53: a str is opened with PIL and converted to RGB, an ndarray is BGR; -> BGR ndarray."""
    if isinstance(f, (str, os.PathLike)):
        from PIL import Image
        return np.ascontiguousarray(np.asarray(Image.open(f).convert("RGB"))[:, :, ::-1])
    a = np.asarray(f)
    if a.ndim == 2:
        a = np.repeat(a[:, :, None], 3, 2)
    return a


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
