"""GPU micro-benchmark of the tcgen05 GEMM over the shapes of the hot path (profiling aid)."""
import ctypes
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisperseg_b200 import _lib  # noqa: E402


def p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def run(lib, M, N, K, gelu, resid, out_f32, bn, iters=10):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    r = out if resid else None
    for _ in range(2):
        _lib.check(lib.wsb_gemm_bf16(p(a), p(w), M, N, K, p(bias), gelu, p(r), p(out), out_f32, bn, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.wsb_gemm_bf16(p(a), p(w), M, N, K, p(bias), gelu, p(r), p(out), out_f32, bn, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9


def main():
    lib = _lib.load()
    M = 120000
    print("encoder shapes, M=%d" % M)
    for name, N, K, gelu, resid, f32 in [("qkv", 3840, 1280, 0, 0, 0), ("out", 1280, 1280, 0, 1, 1), ("fc1", 5120, 1280, 1, 0, 0),
                                         ("fc1-nogelu", 5120, 1280, 0, 0, 0), ("fc2", 1280, 5120, 0, 1, 1),
                                         ("fc2-bf16", 1280, 5120, 0, 0, 0)]:
        for bn in (128, 256):
            ms, tf = run(lib, M, N, K, gelu, resid, f32, bn)
            print("  %-11s N=%5d K=%5d bn=%3d  %8.3f ms  %7.1f TFLOP/s" % (name, N, K, bn, ms, tf))
    # cuBLAS reference (library, for context only)
    a = torch.randn(M, 1280, device="cuda").to(torch.bfloat16)
    w = torch.randn(5120, 1280, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("  cuBLAS fc1-shape (context): %.3f ms %.1f TFLOP/s" % (ms, 2.0 * M * 5120 * 1280 / ms / 1e9))
    print("decode shapes, M=240")
    for name, N, K, gelu, resid, f32 in [("sqkv", 3840, 1280, 0, 0, 0), ("so", 1280, 1280, 0, 1, 1), ("fc1", 5120, 1280, 1, 0, 0),
                                         ("fc2", 1280, 5120, 0, 1, 1)]:
        for bn in (0, 32, 64, 128):
            ms, tf = run(lib, 240, N, K, gelu, resid, f32, bn, iters=50)
            print("  %-5s N=%5d K=%5d bn=%3d  %8.2f us  weights %6.0f GB/s" % (name, N, K, bn, ms * 1000, 2.0 * N * K / ms / 1e6))


if __name__ == "__main__":
    main()
