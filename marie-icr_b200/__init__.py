"""marie-icr_b200 — B200 (sm_100a) implementation of Marie-AI's OCR hot path.

CRAFT text-box detection (reference: marie/boxes/craft_box_processor.py) feeding TrOCR recognition
(reference: marie/document/trocr_ocr_processor.py), behind the reference's BoxProcessor / OcrProcessor
plugin API.  Host code is Python/PyTorch (memory, streams, torch.distributed); all compute is hand-written
CUDA in libmarie_b200.so reached through the C ABI declared in include/marie_b200.h.  There is no CPU
fallback: importing the plugin classes works anywhere, using them requires a B200.
"""
__version__ = "0.1.0"
