"""Kernel-level parity through the C ABI: tcgen05 GEMM (all tile widths, every fused epilogue),
LayerNorm, encoder attention -- against plain torch fp32 on the same bf16-rounded inputs."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


@pytest.fixture(scope="module")
def lib():
    from whisperseg_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 256), (128, 128, 64, 128), (256, 512, 128, 256), (1000, 1280, 1280, 256), (1000, 1280, 1280, 128),
    (777, 384, 1536, 128), (300, 1152, 384, 64), (5, 1536, 384, 32), (240, 5120, 1280, 0), (2500, 3840, 1280, 0),
    (130, 200, 64, 64), (64, 51880, 384, 0),
    (2432 + 70, 16384, 2048, 256),      # B operand 67 MB > half of L2: grouped raster, 16 + 4 m-tiles (short last group)
])
def test_gemm_plain(lib, M, N, K, bn):
    import torch
    from whisperseg_b200 import _lib
    torch.manual_seed(M * 7 + N)
    dev = "cuda"
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    ref = a.float() @ w.float().t() + bias
    out32 = torch.full((M, N), float("nan"), device=dev)
    ldn = N
    if N % 4 == 0:
        _lib.check(lib.wsb_gemm_bf16(_p(a), _p(w), M, N, K, _p(bias), 0, _p(None), _p(out32), 1, bn, None), "gemm f32")
        torch.cuda.synchronize()
        err = (out32 - ref).abs().max().item()
        assert err < 2e-3 * max(1.0, ref.abs().max().item()), "f32 out: max err %g" % err
    if N % 8 == 0:
        out16 = torch.zeros((M, ldn), device=dev, dtype=torch.bfloat16)
        _lib.check(lib.wsb_gemm_bf16(_p(a), _p(w), M, N, K, _p(bias), 1, _p(None), _p(out16), 0, bn, None), "gemm bf16")
        torch.cuda.synchronize()
        refg = torch.nn.functional.gelu(ref)
        err = (out16.float() - refg).abs().max().item()
        assert err < 1e-2 * max(1.0, refg.abs().max().item()), "bf16+gelu out: max err %g" % err


def test_gemm_residual_inplace(lib):
    import torch
    from whisperseg_b200 import _lib
    torch.manual_seed(3)
    M, N, K = 900, 1280, 5120
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    x = torch.randn(M, N, device="cuda")
    ref = x + a.float() @ w.float().t() + bias
    _lib.check(lib.wsb_gemm_bf16(_p(a), _p(w), M, N, K, _p(bias), 0, _p(x), _p(x), 1, 0, None), "gemm resid")
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 3e-3 * ref.abs().max().item()


@pytest.mark.parametrize("rows,d", [(1000, 1280), (37, 384), (500, 512), (3, 768)])
def test_layernorm(lib, rows, d):
    import torch
    from whisperseg_b200 import _lib
    torch.manual_seed(rows)
    x = torch.randn(rows, d, device="cuda") * 3 + 1
    g = torch.randn(d, device="cuda")
    b = torch.randn(d, device="cuda")
    o16 = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    o32 = torch.empty(rows, d, device="cuda")
    _lib.check(lib.wsb_layernorm(_p(x), _p(g), _p(b), _p(o16), _p(o32), rows, d, None), "layernorm")
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5)
    assert (o32 - ref).abs().max().item() < 1e-4
    assert (o16.float() - ref).abs().max().item() < 4e-2


# (the last three give every persistent CTA several (head, window) units: next-unit K / Q / V prefetch, 4, 1 and 3 tiles per unit)
@pytest.mark.parametrize("B,T,H", [(1, 500, 6), (3, 500, 20), (2, 128, 8), (2, 300, 6), (24, 500, 20), (50, 128, 8), (70, 300, 6)])
def test_encoder_attention(lib, B, T, H):
    import torch
    from whisperseg_b200 import _lib
    torch.manual_seed(B * 100 + T)
    d = H * 64
    qkv = torch.randn(B * T, 3 * d, device="cuda")
    qkv[:, :d] *= 0.35                      # q already carries the 1/8 scaling (and then some spread)
    qkv = qkv.to(torch.bfloat16)
    out = torch.zeros(B * T, d, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.wsb_encoder_attention(_p(qkv), _p(out), B, T, H, None), "attention")
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, T, H, 64).transpose(1, 2) for t in qkv.split(d, dim=1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(B * T, d)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, "attention max err %g" % err


@pytest.mark.parametrize("M,N,K", [(16, 3840, 1280), (5, 1280, 1280), (16, 5120, 1280), (9, 1280, 5120), (1, 1152, 384),
                                   (16, 1536, 384), (12, 384, 1536), (3, 52, 96),
                                   (17, 1280, 1280), (32, 3840, 1280), (40, 5120, 1280), (64, 1280, 5120), (64, 3840, 1280),
                                   (50, 1152, 384), (33, 384, 1536), (64, 52, 96)])
@pytest.mark.parametrize("mode", ["ln_f32", "ln_gelu", "bf16_resid", "bf16_f32", "ln_resid", "bf16_gelu", "fold_f32", "fold_gelu"])
def test_gemv16(lib, M, N, K, mode):
    """Skinny linear for <= 64 rows (1, 2 or 4 m-tiles of 16; fused LayerNorm / bias / GELU / residual) against torch fp32 on the
    same bf16-rounded operands.  Tolerance: fp32 accumulation-order noise (2e-3 relative to the output scale),
    bf16 output rounding for the GELU mode."""
    import torch
    from whisperseg_b200 import _lib
    if mode.startswith(("ln", "fold")) and K > 1536:
        pytest.skip("fused LayerNorm is for d_model-wide inputs")
    torch.manual_seed(M * 31 + N + K)
    dev = "cuda"
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    if mode.startswith("fold"):
        # LayerNorm affine folded into the projection (weights.py: fold_layernorm): the kernel reads bf16(x) and
        # applies rstd (acc - mean c1) + c2; reference = fp32 LayerNorm + linear on the unfolded fp32 weights
        x = torch.randn(M, K, device=dev) * 3.0 + 0.5
        gamma, beta = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
        w32 = w.float()
        wf = (w32 * gamma[None, :]).to(torch.bfloat16)
        c1 = wf.float().sum(1).contiguous()
        c2 = (bias + w32 @ beta).contiguous()
        ref = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5) @ w32.t() + bias
        scale = max(1.0, ref.abs().max().item())
        if mode.endswith("f32"):
            out = torch.full((M, N), float("nan"), device=dev)
            _lib.check(lib.wsb_gemv16(_p(x), _p(c1), _p(None), _p(None), _p(wf), _p(c2), M, N, K, 0, _p(out), None), "gemv16 fold")
            torch.cuda.synchronize()
            assert (out - ref).abs().max().item() < 6e-3 * scale      # bf16 rounding of x and of W o gamma
        else:
            out = torch.zeros((M, N), device=dev, dtype=torch.bfloat16)
            _lib.check(lib.wsb_gemv16(_p(x), _p(c1), _p(None), _p(None), _p(wf), _p(c2), M, N, K, 1, _p(out), None), "gemv16 fold")
            torch.cuda.synchronize()
            refg = torch.nn.functional.gelu(ref)
            assert (out.float() - refg).abs().max().item() < 1.2e-2 * scale
        return
    if mode.startswith("ln"):
        x = torch.randn(M, K, device=dev) * 3.0 + 0.5
        gamma, beta = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
        a = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5).to(torch.bfloat16)
        args = (_p(x), _p(gamma), _p(beta), _p(None))
    else:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        args = (_p(None), _p(None), _p(None), _p(a))
    ref = a.float() @ w.float().t() + bias
    scale = max(1.0, ref.abs().max().item())
    if mode.endswith("f32"):
        out = torch.full((M, N), float("nan"), device=dev)
        _lib.check(lib.wsb_gemv16(*args, _p(w), _p(bias), M, N, K, 0, _p(out), None), "gemv16")
        torch.cuda.synchronize()
        assert (out - ref).abs().max().item() < 2e-3 * scale
    elif mode.endswith("gelu"):
        out = torch.zeros((M, N), device=dev, dtype=torch.bfloat16)
        _lib.check(lib.wsb_gemv16(*args, _p(w), _p(bias), M, N, K, 1, _p(out), None), "gemv16")
        torch.cuda.synchronize()
        refg = torch.nn.functional.gelu(ref)
        assert (out.float() - refg).abs().max().item() < 1e-2 * scale
    else:
        resid = torch.randn(M, N, device=dev)
        out = resid.clone()
        _lib.check(lib.wsb_gemv16(*args, _p(w), _p(bias), M, N, K, 2, _p(out), None), "gemv16")
        torch.cuda.synchronize()
        assert (out - (resid + ref)).abs().max().item() < 2e-3 * scale


@pytest.mark.parametrize("M,N,K,splits", [(240, 3840, 1280, 0), (240, 1280, 1280, 0), (240, 5120, 1280, 0), (240, 1280, 5120, 0),
                                          (130, 1280, 1280, 0), (65, 384, 384, 0), (256, 1536, 384, 0), (100, 384, 1536, 0),
                                          (240, 1280, 1280, 1), (240, 1280, 1280, 2), (240, 1280, 1280, 8), (77, 1280, 5120, 8),
                                          (5, 128, 64, 1)])
@pytest.mark.parametrize("mode", ["fold_f32", "fold_gelu", "resid"])
def test_skinny_cluster_linear(lib, M, N, K, splits, mode):
    """Cluster split-K skinny linear (csrc/skinny.cu): DSMEM-reduced partial tiles, folded LayerNorm / in-place
    residual epilogues (+ bf16 copy and per-tile row statistics), against torch fp32.  Tolerances as for gemv16."""
    import torch
    from whisperseg_b200 import _lib
    torch.manual_seed(M * 17 + N + K + splits)
    dev = "cuda"
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    if mode.startswith("fold"):
        x = torch.randn(M, K, device=dev) * 3.0 + 0.5
        gamma, beta = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
        w32 = w.float()
        wf = (w32 * gamma[None, :]).to(torch.bfloat16)
        c1 = wf.float().sum(1).contiguous()
        c2 = (bias + w32 @ beta).contiguous()
        ref = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5) @ w32.t() + bias
        scale = max(1.0, ref.abs().max().item())
        if mode == "fold_f32":
            out = torch.full((M, N), float("nan"), device=dev)
            _lib.check(lib.wsb_skinny_linear(_p(x), _p(c1), _p(None), _p(wf), _p(c2), M, N, K, 0, splits, _p(out), _p(None),
                                             _p(None), None), "skinny fold")
            torch.cuda.synchronize()
            assert (out - ref).abs().max().item() < 6e-3 * scale
        else:
            out = torch.zeros((M, N), device=dev, dtype=torch.bfloat16)
            _lib.check(lib.wsb_skinny_linear(_p(x), _p(c1), _p(None), _p(wf), _p(c2), M, N, K, 1, splits, _p(out), _p(None),
                                             _p(None), None), "skinny fold gelu")
            torch.cuda.synchronize()
            refg = torch.nn.functional.gelu(ref)
            assert (out.float() - refg).abs().max().item() < 1.2e-2 * scale
        return
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    resid = torch.randn(M, N, device=dev)
    ref = resid + a.float() @ w.float().t() + bias
    out = resid.clone()
    xb = torch.zeros((M, N), device=dev, dtype=torch.bfloat16)
    stats = torch.full((N // 128, M, 2), float("nan"), device=dev)
    _lib.check(lib.wsb_skinny_linear(_p(None), _p(None), _p(a), _p(w), _p(bias), M, N, K, 2, splits, _p(out), _p(xb), _p(stats),
                                     None), "skinny resid")
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    assert (out - ref).abs().max().item() < 2e-3 * scale
    assert torch.equal(xb, out.to(torch.bfloat16))
    tiles = out.view(M, N // 128, 128).permute(1, 0, 2)
    assert (stats[..., 0] - tiles.sum(-1)).abs().max().item() < 1e-2 * scale
    assert (stats[..., 1] - (tiles * tiles).sum(-1)).abs().max().item() < 1e-3 * (tiles * tiles).sum(-1).max().item()
