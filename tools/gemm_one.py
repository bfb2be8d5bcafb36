"""Run a few GEMM launches of one shape (target for `ncu --set full`)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisperseg_b200 import _lib  # noqa: E402

lib = _lib.load()
M, N, K, gelu, resid, f32, bn = [int(x) for x in sys.argv[1:8]]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)  # noqa: E731
for _ in range(3):
    _lib.check(lib.wsb_gemm_bf16(p(a), p(w), M, N, K, p(bias), gelu, p(out if resid else None), p(out), f32, bn, None))
torch.cuda.synchronize()
print("done")
