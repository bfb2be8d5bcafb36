"""GPU aid: steady-state time of the skinny-linear kernel (gemv.cu) for the decoder shapes of whisper-large,
by row count."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisperseg_b200 import _lib  # noqa: E402

lib = _lib.load()
shapes = [("qkv  LN->f32 ", 3840, 1280, 0), ("cq   LN->f32 ", 1280, 1280, 0), ("fc1  LN->gelu", 5120, 1280, 1),
          ("so/co  ->res ", 1280, 1280, 2), ("fc2    ->res ", 1280, 5120, 2),
          ("qkv  fold->f32", 3840, 1280, 3), ("cq   fold->f32", 1280, 1280, 3), ("fc1  fold->gelu", 5120, 1280, 4)]
rows = [int(a) for a in sys.argv[1:]] or [16, 32, 48, 64]
for name, N, K, mode in shapes:
    line = []
    for M in rows:
        us = ctypes.c_float()
        _lib.check(lib.wsb_gemv16_bench(M, N, K, mode, 400, 24, ctypes.byref(us)), "gemv16_bench")
        line.append("M=%2d %6.2f us" % (M, us.value))
    print("%s N=%4d K=%4d: %s" % (name, N, K, "   ".join(line)), flush=True)
