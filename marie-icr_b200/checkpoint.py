"""Loads a fairseq TrOCR checkpoint (`trocr-large-printed.pt`, marie/document/trocr_ocr_processor.py:59-70) WITHOUT
fairseq: the file is a pickle of {"model": OrderedDict of tensors, "cfg" / "args": fairseq / omegaconf / argparse
objects, ...}.  `torch.load(weights_only=False)` would import (and execute) whatever classes the pickle names and fails
when fairseq is absent; `weights_only=True` rejects the config objects.  Here a restricted unpickler rebuilds tensors
and plain containers, and turns every other class into an inert attribute bag, so the weights and the few configuration
fields the packer validates (activation_fn, learned positions, ...) can be read and nothing else runs.
"""
import argparse
import collections
import pickle

import torch

_SAFE = {
    ("collections", "OrderedDict"): collections.OrderedDict,
    ("argparse", "Namespace"): argparse.Namespace,
    ("builtins", "set"): set, ("builtins", "frozenset"): frozenset, ("builtins", "list"): list, ("builtins", "dict"): dict,
    ("builtins", "tuple"): tuple, ("builtins", "int"): int, ("builtins", "float"): float, ("builtins", "bool"): bool,
    ("builtins", "str"): str, ("builtins", "bytes"): bytes, ("builtins", "complex"): complex, ("builtins", "slice"): slice,
}
_TORCH_OK = ("torch._utils", "torch", "torch.storage", "torch.serialization", "torch._tensor", "numpy.core.multiarray",
             "numpy._core.multiarray", "numpy", "numpy.dtypes")


class Opaque:
    """stand-in for a class that is not rebuilt (fairseq / omegaconf config objects): keeps the pickled state only"""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        self.__dict__["_state"] = state
        if isinstance(state, dict):
            self.__dict__.update({k: v for k, v in state.items() if isinstance(k, str)})

    def __call__(self, *args, **kwargs):
        return Opaque(*args, **kwargs)

    def __setitem__(self, k, v):
        self.__dict__.setdefault("_items", {})[k] = v

    def append(self, v):
        self.__dict__.setdefault("_list", []).append(v)

    def extend(self, v):
        self.__dict__.setdefault("_list", []).extend(v)


def _opaque_class(module, name):
    return type(name, (Opaque,), {"__module__": "opaque." + module})


class RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _SAFE:
            return _SAFE[(module, name)]
        if module in _TORCH_OK or module.startswith("torch."):
            allowed = (name.startswith("_rebuild") or name.endswith("Storage") or name in ("dtype", "device", "Size", "Tensor", "_reconstruct", "ndarray", "scalar")
                       or name in ("float16", "float32", "bfloat16", "int64", "int32", "uint8", "bool", "float64")
                       or module in ("numpy.dtypes",) or name == "OrderedDict")
            if allowed:
                return super().find_class(module, name)
        return _opaque_class(module, name)


class _PickleModule:
    """what torch.load(pickle_module=...) expects"""
    __name__ = "marie_b200_restricted_pickle"
    Unpickler = RestrictedUnpickler
    load = staticmethod(lambda f, **kw: RestrictedUnpickler(f, **kw).load())


def _find(obj, key, depth=0):
    """first value stored under `key` anywhere inside nested dicts / Namespaces / opaque config objects"""
    if depth > 6 or obj is None:
        return None
    items = None
    if isinstance(obj, dict):
        items = obj
    elif isinstance(obj, (argparse.Namespace, Opaque)):
        items = vars(obj)
    if items is None:
        return None
    if key in items and not isinstance(items[key], (dict, Opaque)):
        v = items[key]
        return getattr(v, "_val", v) if isinstance(v, Opaque) else v
    for k in ("_content", "_state", "model", "cfg", "args"):
        if k in items:
            r = _find(items[k], key, depth + 1)
            if r is not None:
                return r
    for v in items.values():
        if isinstance(v, (dict, argparse.Namespace, Opaque)):
            r = _find(v, key, depth + 1)
            if r is not None:
                return r
    return None


def load_fairseq_checkpoint(path):
    """-> (state_dict, info) where info holds the configuration fields pack_trocr validates (None when the checkpoint
    does not say)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)
    if not isinstance(ckpt, dict) or "model" not in ckpt:
        raise ValueError(f"{path}: not a fairseq checkpoint (no 'model' entry)")
    sd = ckpt["model"]
    if not all(torch.is_tensor(v) for v in sd.values()):
        raise ValueError(f"{path}: 'model' holds non-tensor entries")
    info = {}
    for key in ("activation_fn", "decoder_learned_pos", "decoder_normalize_before", "layernorm_embedding", "no_scale_embedding",
                "deit_arch", "decoder_layers", "decoder_embed_dim", "share_decoder_input_output_embed", "bpe", "arch"):
        for root in (ckpt.get("cfg"), ckpt.get("args")):
            v = _find(root, key)
            if v is not None:
                info[key] = v
                break
    return sd, info
