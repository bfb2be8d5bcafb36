// K4 -- fused encoder self-attention on tcgen05 / TMEM (non-causal, T <= 512 keys, head_dim 64).
//
// Replaces HF WhisperAttention.forward for the encoder (modeling_whisper.py:310-357): softmax(q k^T) v
// with q already scaled (the 1/sqrt(64) factor is folded into the q projection weights at load time).
//
// One CTA per (128-query tile, head, window).  Q, K and V tiles are TMA-loaded straight out of the
// packed [B*T, 3d] QKV activation (128-byte swizzle; keys beyond T are zero-filled by TMA and masked).
//   S = Q K^T      : 2 x (M=128, N=256, K=64) tcgen05.mma into all 512 TMEM columns (fp32)
//   softmax        : thread r owns TMEM lane r = query row r: pass 1 row max, pass 2 exp2 + row sum,
//                    P written as bf16 into a double-buffered, manually 128B-swizzled smem tile
//   O = P V        : per 64-key chunk one (M=128, N=64, K=64) MMA group; V is consumed MN-major
//                    (head_dim contiguous) exactly as TMA delivered it.  O aliases the first 64
//                    columns of S, which are dead once chunk 0 of P has been produced.
//   epilogue       : O / rowsum -> bf16 -> out[B*T, d]
#include "common.cuh"
#include "wsb_internal.h"

namespace wsb {

constexpr int kAttThreads = 128;
constexpr int kAttQ = 128;          // query rows per CTA
constexpr int kAttKeys = 512;       // padded key count
constexpr int kHd = 64;
constexpr int kAttSmem = (kAttQ * kHd + 2 * kAttKeys * kHd + 2 * kAttQ * kHd) * 2 + 1024 + 64;

__global__ void __launch_bounds__(kAttThreads, 1)
encoder_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                         __nv_bfloat16* __restrict__ out, int T, int d) {
    extern __shared__ unsigned char att_smem_raw[];
    unsigned char* smem = att_smem_raw + ((1024u - (smem_u32(att_smem_raw) & 1023u)) & 1023u);
    unsigned char* sQ = smem;                                   // 16 KB
    unsigned char* sK = sQ + kAttQ * kHd * 2;                   // 64 KB
    unsigned char* sV = sK + kAttKeys * kHd * 2;                // 64 KB
    unsigned char* sP = sV + kAttKeys * kHd * 2;                // 2 x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAttQ * kHd * 2);
    uint64_t* bar_load = bars;
    uint64_t* bar_s = bars + 1;
    uint64_t* bar_p = bars + 2;                                 // [2] P buffer consumed by the tensor core
    uint64_t* bar_o = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;

    if (tid == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(&bar_p[0], 1);
        mbar_init(&bar_p[1], 1);
        mbar_init(bar_o, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(bar_load, (kAttQ + 2 * kAttKeys) * kHd * 2);
        tma_load_3d(sQ, &tm_q, bar_load, h * kHd, qt * kAttQ, b);
        tma_load_3d(sK, &tm_kv, bar_load, d + h * kHd, 0, b);
        tma_load_3d(sK + 256 * kHd * 2, &tm_kv, bar_load, d + h * kHd, 256, b);
        tma_load_3d(sV, &tm_kv, bar_load, 2 * d + h * kHd, 0, b);
        tma_load_3d(sV + 256 * kHd * 2, &tm_kv, bar_load, 2 * d + h * kHd, 256, b);
        mbar_wait(bar_load, 0);
        tc_fence_after();
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
        const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + half * 256 * kHd * 2));
#pragma unroll
            for (int k = 0; k < kHd / 16; ++k)
                umma_bf16_ss(tmem + half * 256, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
    }
    mbar_wait(bar_s, 0);
    tc_fence_after();

    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;

    // pass 1: row max over the T valid keys
    float rmax = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < kAttKeys / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(lane_base + c * 32, r);
        tmem_ld_wait();
        const int n0 = c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (n0 + i < T) rmax = fmaxf(rmax, __uint_as_float(r[i]));
    }
    const float mscaled = rmax * kLog2e;

    // pass 2: P chunks + PV MMAs
    float rsum = 0.0f;
    const int row = tid;                                   // query row within the tile == TMEM lane
#pragma unroll 1
    for (int c = 0; c < kAttKeys / 64; ++c) {
        const int buf = c & 1;
        if (c >= 2) mbar_wait(&bar_p[buf], ((c >> 1) - 1) & 1);
        unsigned char* pbuf = sP + buf * kAttQ * kHd * 2 + row * 128;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t r[32];
            tmem_ld_32x32(lane_base + c * 64 + hlf * 32, r);
            tmem_ld_wait();
            const int n0 = c * 64 + hlf * 32;
            float pv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float e = (n0 + i < T) ? exp2f(fmaf(__uint_as_float(r[i]), kLog2e, -mscaled)) : 0.0f;
                const float eb = __bfloat162float(__float2bfloat16(e));
                pv[i] = eb;
                rsum += eb;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {                  // four 16-byte chunks of this 32-key half
                uint4 pk;
                pk.x = pack_bf16x2(pv[8 * j + 0], pv[8 * j + 1]);
                pk.y = pack_bf16x2(pv[8 * j + 2], pv[8 * j + 3]);
                pk.z = pack_bf16x2(pv[8 * j + 4], pv[8 * j + 5]);
                pk.w = pack_bf16x2(pv[8 * j + 6], pv[8 * j + 7]);
                const int chunk = hlf * 4 + j;             // 16-byte chunk index within the 128-byte row
                *reinterpret_cast<uint4*>(pbuf + ((chunk ^ (row & 7)) << 4)) = pk;
            }
        }
        fence_proxy_async_smem();                          // generic-proxy smem writes -> visible to the MMA
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);     // B (= V) is MN-major
            const uint64_t dp = umma_desc_k_sw128(smem_u32(sP + buf * kAttQ * kHd * 2));
            const uint64_t dv = umma_desc_mn_sw128(smem_u32(sV + c * 64 * kHd * 2));
#pragma unroll
            for (int k = 0; k < 4; ++k)                    // 16 keys per MMA: P advances 32 B, V advances 16 rows
                umma_bf16_ss(tmem, dp + 2 * k, dv + ((16 * kHd * 2) >> 4) * k, idesc_o, (c > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bar_p[buf]);
            if (c == kAttKeys / 64 - 1) umma_commit(bar_o);
        }
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();

    // epilogue: O / rowsum
    const int q_row = qt * kAttQ + row;
    const float inv = 1.0f / rsum;
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t r[32];
        tmem_ld_32x32(lane_base + hlf * 32, r);
        tmem_ld_wait();
        if (q_row < T) {
            __nv_bfloat16* o = out + (static_cast<size_t>(b) * T + q_row) * d + h * kHd + hlf * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 pk;
                pk.x = pack_bf16x2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
                pk.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
                pk.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
                pk.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
                *reinterpret_cast<uint4*>(o + i) = pk;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

int encoder_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int n_heads, cudaStream_t stream) {
    WSB_REQUIRE(T <= kAttKeys && T > 0, "encoder attention supports up to 512 positions");
    if (B <= 0) return 0;
    const int d = n_heads * kHd;
    static bool attr_set = false;
    if (!attr_set) {
        WSB_CHECK_CUDA(cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
        attr_set = true;
    }
    CUtensorMap tm_q, tm_kv;
    uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(T) * 3 * d * 2};
    uint32_t box_q[3] = {static_cast<uint32_t>(kHd), static_cast<uint32_t>(kAttQ), 1};
    uint32_t box_kv[3] = {static_cast<uint32_t>(kHd), 256, 1};
    int rc = make_tmap_bf16(&tm_q, qkv, 3, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_bf16(&tm_kv, qkv, 3, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    dim3 grid(ceil_div(T, kAttQ), n_heads, B);
    encoder_attention_kernel<<<grid, kAttThreads, kAttSmem, stream>>>(tm_q, tm_kv, out, T, d);
    WSB_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace wsb
