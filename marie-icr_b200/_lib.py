"""ctypes binding of libmarie_b200.so (C ABI in include/marie_b200.h).

Fails loudly when the shared object is missing or no sm_100 device is present — the product path never
falls back to a CPU implementation.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmarie_b200.so")

_lib = None
_lock = threading.Lock()

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float


class MarieB200Error(RuntimeError):
    pass


def load_library():
    """Loads (once) and returns the ctypes handle. Raises if the extension has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MarieB200Error(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                    "There is no CPU fallback for this path.")
            lib = ctypes.CDLL(LIB_PATH)
            lib.mb_version.restype = ctypes.c_char_p
            lib.mb_last_error.restype = ctypes.c_char_p
            lib.mb_last_error.argtypes = [c_void_p]
            lib.mb_launch_count.restype = ctypes.c_ulonglong
            lib.mb_launch_count.argtypes = [c_void_p]
            lib.mb_init.argtypes = [c_int, ctypes.POINTER(c_void_p)]
            lib.mb_free.argtypes = [c_void_p]
            lib.mb_free.restype = None
            _lib = lib
    return _lib


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


class _CurrentStream:
    """Placeholder argument: Context.call replaces it by torch's current stream OF THE CONTEXT'S DEVICE (not of whatever
    device happens to be current in the calling thread)."""


_CUR_STREAM = _CurrentStream()


def cur_stream():
    return _CUR_STREAM


class Context:
    """One mb_ctx per (process, device)."""

    _instances = {}

    def __init__(self, device=0):
        self.lib = load_library()
        h = c_void_p()
        rc = self.lib.mb_init(int(device), ctypes.byref(h))
        if rc != 0:
            raise MarieB200Error(
                f"mb_init(device={device}) failed with code {rc}: a CUDA sm_100 (B200) device is required; "
                "there is no CPU fallback")
        self.handle = h
        self.device = int(device)

    @classmethod
    def get(cls, device=0):
        device = int(device)
        if device not in cls._instances:
            cls._instances[device] = Context(device)
        return cls._instances[device]

    def check(self, rc, what=""):
        if rc != 0:
            msg = self.lib.mb_last_error(self.handle)
            raise MarieB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def call(self, name, *args):
        fn = getattr(self.lib, name)
        fn.restype = c_int
        if any(a is _CUR_STREAM for a in args):
            import torch
            stream = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            args = tuple(stream if a is _CUR_STREAM else a for a in args)
        self.check(fn(self.handle, *args), name)

    @property
    def torch_dtype(self):
        """torch dtype of the library's 16-bit buffers (mb_get_dtype): float16 (default) or bfloat16."""
        import torch
        self.lib.mb_get_dtype.argtypes = [c_void_p]
        return torch.float16 if self.lib.mb_get_dtype(self.handle) == 1 else torch.bfloat16

    def set_dtype(self, dtype):
        """dtype: torch.float16 / torch.bfloat16 (or 'fp16' / 'bf16').  Only before weights are loaded."""
        code = 1 if str(dtype) in ("torch.float16", "fp16", "f16", "float16") else 0
        self.call("mb_set_dtype", c_int(code))

    def profile(self, enable):
        self.call("mb_profile_enable", c_int(1 if enable else 0))

    def profile_read(self):
        out = (ctypes.c_double * 3)()
        self.call("mb_profile_read", out)
        return dict(ms=out[0], flops=out[1], launches=int(out[2]))

    @property
    def launches(self):
        return int(self.lib.mb_launch_count(self.handle))

    def close(self):
        if self.handle:
            self.lib.mb_free(self.handle)
            self.handle = None
            Context._instances.pop(self.device, None)
