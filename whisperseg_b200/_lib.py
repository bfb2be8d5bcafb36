"""ctypes binding of libwsb.so (the C ABI in include/wsb.h).  No CPU fallback: if the library is
missing or there is no sm_100 device, the product path raises."""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WSB_LIB") or os.path.join(HERE, "libwsb.so")

_lib = None
_lock = threading.Lock()


class WsbError(RuntimeError):
    pass


class ModelConfig(ctypes.Structure):
    _fields_ = [("d_model", ctypes.c_int), ("n_heads", ctypes.c_int), ("n_layers", ctypes.c_int),
                ("ffn_dim", ctypes.c_int), ("vocab_size", ctypes.c_int), ("n_mels", ctypes.c_int),
                ("n_cols", ctypes.c_int), ("max_target_positions", ctypes.c_int), ("max_batch", ctypes.c_int)]


EXPORTS = {
    # name: (restype, argtypes)
    "wsb_abi_version": (ctypes.c_int, []),
    "wsb_last_error": (ctypes.c_char_p, []),
    "wsb_launch_count": (ctypes.c_longlong, [ctypes.c_int]),
    "wsb_set_sm_reserve": (ctypes.c_int, [ctypes.c_int]),
    "wsb_logmel_plan_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "wsb_logmel_plan_destroy": (None, [ctypes.c_void_p]),
    "wsb_logmel_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "wsb_model_create": (ctypes.c_int, [ctypes.POINTER(ModelConfig), ctypes.POINTER(ctypes.c_char_p),
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "wsb_model_destroy": (None, [ctypes.c_void_p]),
    "wsb_model_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p]),
    "wsb_workspace_bytes_for": (ctypes.c_size_t, [ctypes.POINTER(ModelConfig)]),
    "wsb_model_fold_fallback": (ctypes.c_int, [ctypes.c_void_p]),
    "wsb_mega_trace": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]),
    "wsb_encode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "wsb_generate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_void_p]),
    "wsb_generate_beam": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32),
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                         ctypes.c_void_p]),
    "wsb_beam_selftest": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wsb_gemv16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p]),
    "wsb_skinny_linear": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "wsb_gemv16_bench": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_float)]),
    "wsb_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "wsb_profile_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong),
                                        ctypes.POINTER(ctypes.c_double)]),
    "wsb_gemm_bf16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_void_p]),
    "wsb_layernorm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "wsb_encoder_attention": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p]),
}


def load():
    """Load (building on first use if nvcc is present) and type the library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        # (re)build when the sources changed: build() compares a digest of csrc/ + include/ with the stamp next to
        # the library and is a no-op when they match.  Without nvcc (a deployment box) the shipped .so is used as is.
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if "WSB_LIB" not in os.environ and (os.path.isfile(nvcc) or not os.path.isfile(LIB_PATH)):
            try:
                from .build import build
                build()
            except Exception as e:  # noqa: BLE001
                if not os.path.isfile(LIB_PATH):
                    raise WsbError("libwsb.so is not built and could not be built here (%s); run "
                                   "`python -m whisperseg_b200.build`" % e) from e
                raise WsbError("libwsb.so is out of date and rebuilding it failed (%s)" % e) from e
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)            # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if lib.wsb_abi_version() != 1:
            raise WsbError("libwsb.so ABI version mismatch")
        _lib = lib
        return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().wsb_last_error()
        raise WsbError("%s failed (status %d): %s" % (what or "libwsb call", rc, (msg or b"").decode(errors="replace")))
