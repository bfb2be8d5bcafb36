"""Run the encoder attention kernel and the decode cross-attention path a few times (ncu targets)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisperseg_b200 import _lib  # noqa: E402

lib = _lib.load()
B, T, H = 240, 500, 20
d = H * 64
qkv = (torch.randn(B * T, 3 * d, device="cuda") * 0.5).to(torch.bfloat16)
out = torch.zeros(B * T, d, device="cuda", dtype=torch.bfloat16)
p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
for _ in range(3):
    _lib.check(lib.wsb_encoder_attention(p(qkv), p(out), B, T, H, None))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.wsb_encoder_attention(p(qkv), p(out), B, T, H, None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("encoder attention B=%d: %.3f ms  %.1f TFLOP/s" % (B, ms, 4.0 * B * H * T * T * 64 / ms / 1e9))
