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
    if _cache:
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
