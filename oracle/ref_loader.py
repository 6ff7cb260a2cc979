"""Imports the reference's own CRAFT / line-merge modules by path (container only: /root/reference is absent
on the GPU box).  Follows SURVEY.md Appendix C.  Used to validate the restatements and to generate goldens."""
import importlib.util
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("MARIE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "marie", "models", "craft"))


_cache = {}


def load():
    """Returns a dict with the reference modules: craft (CRAFT class), craft_utils, imgproc, overlap, lines."""
    if "craft" in _cache:
        return _cache
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    os.makedirs("/tmp/fragments", exist_ok=True)   # debug imwrite targets inside the reference code
    for name in ("skimage", "skimage.io"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    craft_dir = os.path.join(REF_ROOT, "marie", "models", "craft")
    if craft_dir not in sys.path:
        sys.path.insert(0, craft_dir)
    import craft as ref_craft          # noqa: E402  (reference file marie/models/craft/craft.py)
    import craft_utils as ref_craft_utils  # noqa: E402
    import imgproc as ref_imgproc      # noqa: E402

    for pkg in ("marie", "marie.logging_core", "marie.utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    if "marie.logging_core.predefined" not in sys.modules:
        pre = types.ModuleType("marie.logging_core.predefined")
        lg = logging.getLogger("oracle.ref")
        lg.setLevel(logging.ERROR)
        pre.default_logger = lg
        sys.modules["marie.logging_core.predefined"] = pre

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    overlap = _load("marie.utils.overlap", "marie/utils/overlap.py")
    lines = _load("ref_line_processor", "marie/boxes/line_processor.py")
    _cache.update(craft=ref_craft, craft_utils=ref_craft_utils, imgproc=ref_imgproc, overlap=overlap,
                  lines=lines)
    return _cache


def load_ocr_processor():
    """The reference's OcrProcessor base class (marie/document/ocr_processor.py:34-267), loaded by path behind stub
    modules for its logging / font / filesystem helpers.  Its `recognize` is the result-assembly oracle."""
    if "ocr_processor" in _cache:
        return _cache["ocr_processor"]
    load()

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("marie.registry_base", "marie/registry_base.py")
    _load("marie.base_handler", "marie/base_handler.py")
    dt = types.ModuleType("marie.utils.draw_truetype")
    dt.determine_font_size = lambda h: 10
    dt.get_default_font = lambda s: None
    sys.modules["marie.utils.draw_truetype"] = dt
    ut = types.ModuleType("marie.utils.utils")

    def ensure_exists(d):
        os.makedirs(d, exist_ok=True)
        return d
    ut.ensure_exists = ensure_exists
    sys.modules["marie.utils.utils"] = ut
    for pkg in ("marie.document",):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    mod = _load("ref_ocr_processor", "marie/document/ocr_processor.py")
    _cache["ocr_processor"] = mod.OcrProcessor
    return mod.OcrProcessor


def load_image_utils():
    """marie/utils/image_utils.py (crop_to_content, ensure_max_page_size, hash_frames_fast) loaded by path behind a stub
    for marie.timer."""
    if "image_utils" in _cache:
        return _cache["image_utils"]
    load()
    if "marie.timer" not in sys.modules:
        t = types.ModuleType("marie.timer")
        t.Timer = object
        sys.modules["marie.timer"] = t
    spec = importlib.util.spec_from_file_location("ref_image_utils", os.path.join(REF_ROOT, "marie/utils/image_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache["image_utils"] = mod
    return mod


def load_box_loop():
    """Source-level extraction (ast) from marie/boxes/craft_box_processor.py, which cannot be imported (it pulls the
    whole `marie` package): the module-level `crop_poly_low` (:42-73) and the per-box loop of
    `BoxProcessorCraft.extract_bounding_boxes` (:499-537: int32 truncation, boundingRect, +4 px expansion, crop,
    find_line_number, debug jpg).  Returns run(image, bboxes, lines_bboxes, crops_dir) -> (rects, fragments, line_numbers)
    that executes the reference's own statements."""
    if "box_loop" in _cache:
        return _cache["box_loop"]
    import ast
    import cv2
    import numpy as np
    mods = load()
    path = os.path.join(REF_ROOT, "marie", "boxes", "craft_box_processor.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "crop_poly_low")
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BoxProcessorCraft")
    method = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "extract_bounding_boxes")
    loop = next(n for n in ast.walk(method) if isinstance(n, ast.For) and isinstance(n.iter, ast.Call)
                and getattr(n.iter.func, "id", "") == "enumerate" and getattr(n.iter.args[0], "id", "") == "bboxes")
    code_fn = compile(ast.Module(body=[fn], type_ignores=[]), path, "exec")
    code_loop = compile(ast.Module(body=[loop], type_ignores=[]), path, "exec")

    def run(image, bboxes, lines_bboxes=(), crops_dir="/tmp/fragments"):
        ns = dict(cv2=cv2, np=np, os=os)
        exec(code_fn, ns)
        os.makedirs(crops_dir, exist_ok=True)
        ns.update(image=image, bboxes=bboxes, lines_bboxes=list(lines_bboxes), crops_dir=crops_dir, ms=0,
                  max_h=image.shape[0], max_w=image.shape[1], rect_from_poly=[], rect_line_numbers=[], fragments=[],
                  find_line_number=mods["lines"].find_line_number, paste_fragment=lambda *a, **k: None, pil_image=None)
        exec(code_loop, ns)
        return ns["rect_from_poly"], ns["fragments"], ns["rect_line_numbers"]

    run.crop_poly_low = lambda img, poly: (lambda ns: (exec(code_fn, ns), ns["crop_poly_low"](img, poly))[1])(dict(cv2=cv2, np=np))
    _cache["box_loop"] = run
    return run


def load_trocr_deit():
    """The reference's TrOCR encoder class, AdaptedVisionTransformer (marie/models/unilm/trocr/deit.py:59-146) and its
    beit_*_patch16_384 factories (:323-337), executed from the reference file.  Its base class lives in timm==0.6.12
    (absent), so a stub `timm` package supplies it from REFERENCE-HELD code: Mlp / Attention / Block / PatchEmbed of the
    vendored copy of timm's ViT in marie/boxes/dit/ditod/deit.py:44-167.  Only the constructor plumbing of timm's
    VisionTransformer (which parameters exist and what they are called) is restated here; forward_features, the
    blocks' arithmetic and the factory arguments (qkv_bias=False, LayerNorm eps 1e-6, depth / width / heads) are the
    reference's own statements."""
    if "trocr_deit" in _cache:
        return _cache["trocr_deit"]
    import torch
    import torch.nn as nn
    sys.dont_write_bytecode = True
    layers = types.ModuleType("timm.models.layers")
    layers.drop_path = lambda x, p=0.0, training=False: x                     # eval mode / p = 0: identity
    layers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    layers.trunc_normal_ = nn.init.trunc_normal_
    timm = types.ModuleType("timm")
    timm.__path__ = []
    models = types.ModuleType("timm.models")
    models.__path__ = []
    models.register_model = lambda f: f
    models.layers = layers
    timm.models = models
    saved = {k: sys.modules.get(k) for k in ("timm", "timm.models", "timm.models.layers", "timm.models.vision_transformer")}
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    try:
        ditod = _load("ref_ditod_deit", "marie/boxes/dit/ditod/deit.py")

        class PatchEmbedFlat(ditod.PatchEmbed):
            """timm 0.6.12 PatchEmbed.forward = proj -> flatten(2).transpose(1, 2) (the ditod copy leaves the map 2-D)"""

            def forward(self, x):
                return super().forward(x).flatten(2).transpose(1, 2)

        class VisionTransformer(nn.Module):
            """constructor plumbing of timm 0.6.12 VisionTransformer: parameter / sub-module names only"""

            def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                         qkv_bias=True, norm_layer=None, drop_rate=0.0, **kw):
                super().__init__()
                norm_layer = norm_layer or nn.LayerNorm
                self.embed_dim, self.num_tokens = embed_dim, 1
                self.patch_embed = PatchEmbedFlat(img_size, patch_size, in_chans, embed_dim)
                self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
                self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
                self.pos_drop = nn.Dropout(drop_rate)
                self.blocks = nn.Sequential(*[ditod.Block(embed_dim, num_heads, mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer)
                                              for _ in range(depth)])
                self.norm = norm_layer(embed_dim)

            def init_weights(self, mode=""):
                pass

        vt = types.ModuleType("timm.models.vision_transformer")
        vt.VisionTransformer, vt._cfg, vt.Attention, vt.Block = VisionTransformer, ditod._cfg, ditod.Attention, ditod.Block
        sys.modules["timm.models.vision_transformer"] = vt
        mod = _load("ref_trocr_deit", "marie/models/unilm/trocr/deit.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["trocr_deit"] = mod
    return mod


def load_generator():
    """The reference's own search loop, TextRecognitionGenerator._generate (marie/models/unilm/trocr/generator.py:11-374),
    executed from the reference file.  Its base class is fairseq's SequenceGenerator (absent: unpinned git HEAD), so a
    stub `fairseq.sequence_generator` supplies the three fairseq pieces the loop calls — BeamSearch.step,
    SequenceGenerator.finalize_hypos / is_finished — restated from fairseq's published algorithm, and the
    attributes `build_generator` sets (task.py:165-276: beam, max_len_b 200, min_len 1, normalize_scores, len_penalty 1,
    unk_penalty 0, temperature 1).  Returns (TextRecognitionGenerator, make_model) where make_model(sd, cfg) adapts the
    oracle's encoder / incremental decoder to the EnsembleModel interface the loop uses."""
    if "generator" in _cache:
        return _cache["generator"]
    import math
    import torch
    import torch.nn.functional as F
    sys.dont_write_bytecode = True

    class BeamSearch:
        """fairseq.search.BeamSearch"""
        supports_constraints = False
        stop_on_max_len = False

        def init_constraints(self, batch_constraints, beam_size):
            pass

        def prune_sentences(self, batch_idxs):
            pass

        def update_constraints(self, active_hypos):
            pass

        def step(self, step, lprobs, scores, prev_output_tokens=None, original_batch_idxs=None):
            bsz, beam_size, vocab_size = lprobs.size()
            if step == 0:
                lprobs = lprobs[:, ::beam_size, :].contiguous()        # all beams are identical at the first step
            else:
                lprobs = lprobs + scores[:, :, step - 1].unsqueeze(-1)
            flat = lprobs.view(bsz, -1)
            scores_buf, indices_buf = torch.topk(flat, k=min(beam_size * 2, flat.size(1) - 1))
            beams_buf = torch.div(indices_buf, vocab_size, rounding_mode="trunc")
            return scores_buf, indices_buf.fmod(vocab_size), beams_buf

    class SequenceGenerator:
        """attribute set of fairseq.sequence_generator.SequenceGenerator.__init__ + finalize_hypos / is_finished"""

        def __init__(self, model, vocab_size, beam_size=1, max_len_a=0, max_len_b=200, min_len=1, normalize_scores=True,
                     len_penalty=1.0, unk_penalty=0.0, temperature=1.0, pad=1, unk=3, eos=2):
            self.model, self.vocab_size, self.beam_size = model, vocab_size, beam_size
            self.max_len_a, self.max_len_b, self.min_len = max_len_a, max_len_b, min_len
            self.normalize_scores, self.len_penalty, self.unk_penalty = normalize_scores, len_penalty, unk_penalty
            self.temperature, self.pad, self.unk, self.eos = temperature, pad, unk, eos
            self.match_source_len, self.lm_model, self.repeat_ngram_blocker = False, None, None
            self.should_set_src_lengths = False
            self.search = BeamSearch()

        def is_finished(self, step, unfin_idx, max_len, finalized_sent_len, beam_size):
            assert finalized_sent_len <= beam_size
            return finalized_sent_len == beam_size or step == max_len

        def finalize_hypos(self, step, bbsz_idx, eos_scores, tokens, scores, finalized, finished, beam_size, attn,
                           src_lengths, max_len):
            assert bbsz_idx.numel() == eos_scores.numel()
            tokens_clone = tokens.index_select(0, bbsz_idx)[:, 1:step + 2]
            tokens_clone[:, step] = self.eos
            pos_scores = scores.index_select(0, bbsz_idx)[:, :step + 1]
            pos_scores[:, step] = eos_scores
            pos_scores[:, 1:] = pos_scores[:, 1:] - pos_scores[:, :-1]
            if self.normalize_scores:
                eos_scores /= (step + 1) ** self.len_penalty
            cum_unfin, prev = [], 0
            for f in finished:
                if f:
                    prev += 1
                else:
                    cum_unfin.append(prev)
            cum_fin = torch.tensor(cum_unfin, dtype=torch.int).to(bbsz_idx)
            unfin_idx = torch.div(bbsz_idx, beam_size, rounding_mode="trunc")
            sent = unfin_idx + torch.index_select(cum_fin, 0, unfin_idx)
            seen = (sent << 32) + unfin_idx
            unique_seen = torch.unique(seen).tolist()
            sent_list = sent.tolist()
            for i in range(bbsz_idx.size(0)):
                if len(finalized[sent_list[i]]) < beam_size:
                    finalized[sent_list[i]].append({"tokens": tokens_clone[i], "score": eos_scores[i], "attention": torch.empty(0),
                                                    "alignment": torch.empty(0), "positional_scores": pos_scores[i]})
            newly_finished = []
            for unique_s in unique_seen:
                unique_sent = unique_s >> 32
                unique_unfin_idx = unique_s - (unique_sent << 32)
                if not finished[unique_sent] and self.is_finished(step, unique_unfin_idx, max_len, len(finalized[unique_sent]),
                                                                  beam_size):
                    finished[unique_sent] = True
                    newly_finished.append(unique_unfin_idx)
            return newly_finished

    fs = types.ModuleType("fairseq")
    fs.__path__ = []
    sg = types.ModuleType("fairseq.sequence_generator")
    sg.SequenceGenerator = SequenceGenerator
    saved = {k: sys.modules.get(k) for k in ("fairseq", "fairseq.sequence_generator")}
    sys.modules.update({"fairseq": fs, "fairseq.sequence_generator": sg})
    try:
        spec = importlib.util.spec_from_file_location("ref_trocr_generator",
                                                      os.path.join(REF_ROOT, "marie/models/unilm/trocr/generator.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    def make_model(sd, cfg):
        from oracle import trocr

        class Model:
            """fairseq EnsembleModel interface (one model) over the oracle's encoder / incremental decoder"""
            models_size = 1

            def __init__(self):
                self.st = None

            def forward_encoder(self, net_input):
                enc = trocr.encoder_forward(sd, cfg, net_input["imgs"])               # trocr_models.py:508-524
                return [{"encoder_out": [enc.transpose(0, 1)], "encoder_padding_mask": [torch.zeros(enc.shape[0], enc.shape[1])]}]

            def max_decoder_positions(self):
                return cfg.max_positions

            def reorder_encoder_out(self, encoder_outs, new_order):
                e = encoder_outs[0]
                return [{"encoder_out": [e["encoder_out"][0].index_select(1, new_order)],
                         "encoder_padding_mask": [e["encoder_padding_mask"][0].index_select(0, new_order)]}]

            def reorder_incremental_state(self, incremental_states, new_order):
                if self.st is not None:
                    self.st.reorder(new_order)
                    self.st.cross = [(k[new_order], v[new_order]) for k, v in self.st.cross]

            def forward_decoder(self, tokens, encoder_outs, incremental_states, temperature=1.0):
                if self.st is None:
                    self.st = trocr.DecoderState(sd, cfg, encoder_outs[0]["encoder_out"][0].transpose(0, 1), beam=1)
                logits = self.st.step(tokens[:, -1], tokens.shape[1] - 1)
                return F.log_softmax(logits.float() / temperature, -1), None
        return Model()

    _cache["generator"] = (mod.TextRecognitionGenerator, make_model)
    return _cache["generator"]


def load_bpe():
    """GPT2BPEEnhancedSpace (marie/models/unilm/trocr/bpe.py:10-67) executed from the reference file; its base class
    fairseq GPT2BPE is absent, so a stub supplies `self.bpe` = the GPT-2 byte-level decoder restated from fairseq's
    gpt2_bpe_utils.Encoder.decode (join vocabulary strings, map characters back to bytes, UTF-8 decode with
    errors='replace').  Returns make(encoder_json_path) -> object with the reference's .decode(str)."""
    if "bpe" in _cache:
        return _cache["bpe"]
    import json
    sys.dont_write_bytecode = True

    def bytes_to_unicode():
        bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
        cs, n = bs[:], 0
        for b in range(256):
            if b not in bs:
                bs.append(b)
                cs.append(256 + n)
                n += 1
        return dict(zip(bs, [chr(c) for c in cs]))

    class Encoder:
        def __init__(self, encoder):
            self.decoder = {v: k for k, v in encoder.items()}
            self.byte_decoder = {v: k for k, v in bytes_to_unicode().items()}

        def decode(self, tokens):
            text = "".join([self.decoder.get(token, token) for token in tokens])
            return bytearray([self.byte_decoder[c] for c in text]).decode("utf-8", errors="replace")

    class GPT2BPE:
        def __init__(self, cfg):
            with open(cfg.gpt2_encoder_json, encoding="utf-8") as f:
                self.bpe = Encoder(json.load(f))

    names = ("fairseq", "fairseq.data", "fairseq.data.encoders", "fairseq.data.encoders.gpt2_bpe")
    saved = {k: sys.modules.get(k) for k in names}
    for k in names:
        m = types.ModuleType(k)
        m.__path__ = []
        sys.modules[k] = m
    sys.modules["fairseq.data.encoders"].register_bpe = lambda name, dataclass=None: (lambda cls: cls)
    sys.modules["fairseq.data.encoders.gpt2_bpe"].GPT2BPE = GPT2BPE
    sys.modules["fairseq.data.encoders.gpt2_bpe"].GPT2BPEConfig = object
    try:
        spec = importlib.util.spec_from_file_location("ref_trocr_bpe", os.path.join(REF_ROOT, "marie/models/unilm/trocr/bpe.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    def make(encoder_json_path):
        return mod.GPT2BPEEnhancedSpace(types.SimpleNamespace(gpt2_encoder_json=encoder_json_path))

    _cache["bpe"] = make
    return make


def load_region_loop():
    """OcrEngine.__process_extract_regions (marie/ocr/ocr_engine.py:223-414) extracted from the source (the module pulls
    the whole `marie` package) and executed as a plain function over stand-in `self` / processors.  Returns
    run(frames, regions, pms_mode, box_processor, icr_processor, PSMode) -> {"regions": [...], "extended": [...]}."""
    if "region_loop" in _cache:
        return _cache["region_loop"]
    import ast
    from itertools import chain
    import numpy as np
    iu = load_image_utils()
    path = os.path.join(REF_ROOT, "marie", "ocr", "ocr_engine.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "OcrEngine")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name.endswith("__process_extract_regions"))
    fn.name = "process_extract_regions"
    for a in fn.args.args:                       # `box_processor: BoxProcessor` / `icr_processor: OcrProcessor`
        a.annotation = None
    code = compile(ast.Module(body=[fn], type_ignores=[]), path, "exec")

    def run(frames, regions, pms_mode, box_processor, icr_processor, PSMode):
        ns = dict(np=np, chain=chain, hash_frames_fast=iu.hash_frames_fast, crop_to_content=iu.crop_to_content, PSMode=PSMode,
                  bbox_cache={}, encodeToBase64=lambda im: "")
        exec(code, ns)
        me = types.SimpleNamespace(logger=logging.getLogger("oracle.ref"))
        return ns["process_extract_regions"](me, frames, "q", "c", pms_mode, regions, box_processor, icr_processor)

    _cache["region_loop"] = run
    return run
