// HBM-bound kernels of the training step (fwd -> CE over text positions -> bwd -> grad clip
// -> AdamW/Lion).  They replace what autograd + torch.optim do around the reference's modules:
//   kx_act_layernorm_fwd   ffn_layernorm(gelu(u)) as one pass (training keeps u, the pre-activation)
//   kx_layernorm_bwd       nn.LayerNorm backward (+ GELU backward, + residual-gradient add, + bf16 copy), with
//                          per-CTA partial d(gamma), d(beta) and column sums, folded by kx_colpartials_reduce
//   kx_colsum_bf16         bias gradients
//   kx_xpos_bwd            transpose of the xPos rotation on dq / dk
//   kx_ce_fwd_bwd          softmax cross-entropy over the text rows of the spliced sequence, d(logits) in bf16
//   kx_embed_bwd           scatter-add of d(x0) into the token-embedding and position tables
//   kx_sumsq / kx_clip_scale / kx_adamw_step / kx_lion_step   gradient-norm clipping and the fused optimizers
// Everything is fp32 math on 16-byte vector accesses; one pass over each operand.
#include "kx_internal.h"
#include "ptx.cuh"
#include "philox.cuh"
#include <mutex>

namespace kx {

constexpr int ROW_THREADS = 256;

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float2 f = __bfloat1622float2(h[u]);
        v[2 * u] = f.x; v[2 * u + 1] = f.y;
    }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = q;
}

// sum of (a, b) over the 256 threads of the CTA, result broadcast to every thread.  `red` is 2 x 8 floats of smem;
// two alternating halves would be needed for back-to-back calls, so every call is followed by a __syncthreads().
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[w] = a; red[8 + w] = b; }
    __syncthreads();
    a = 0.f; b = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_THREADS / 32; ++i) { a += red[i]; b += red[8 + i]; }
    __syncthreads();
}

// sums of four values over the CTA in one exchange (one __syncthreads pair): red holds 4 x 8 floats
__device__ __forceinline__ void block_sum4(float& a, float& b, float& c, float& d, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[w] = a; red[8 + w] = b; red[16 + w] = c; red[24 + w] = d; }
    __syncthreads();
    a = 0.f; b = 0.f; c = 0.f; d = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_THREADS / 32; ++i) { a += red[i]; b += red[8 + i]; c += red[16 + i]; d += red[24 + i]; }
    __syncthreads();
}

// Phi(x) = 0.5 * (1 + erf(x / sqrt2)) with ONE MUFU: erfc(a) = 2^(-Q(a)), a = |x| / sqrt2, Q the degree-5 fit of
// ptx.cuh's gelu_erf_x2 (max |erf error| 8.9e-7).  erff() costs ~25 FP32 instructions per element, which made the
// GELU kernels issue-bound instead of HBM-bound.
__device__ __forceinline__ float gelu_cdf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float q = fmaf(z, -0.003024620935320854f, 0.029882797971367836f);
    q = fmaf(q, z, -0.14901681244373322f);
    q = fmaf(q, z, -0.9183504581451416f);
    q = fmaf(q, z, -1.6279104948043823f);
    const float e = 0.5f * ex2_approx(q * z);               // 0.5 * erfc(|x| / sqrt2)
    return x >= 0.f ? 1.0f - e : e;
}
__device__ __forceinline__ float gelu_exact(float x) { return x * gelu_cdf(x); }
__device__ __forceinline__ float gelu_grad(float x) {       // Phi(x) + x * phi(x)
    const float pdf = 0.3989422804014327f * ex2_approx(-0.72134752044448170368f * x * x);
    return fmaf(x, pdf, gelu_cdf(x));
}

// ----------------------------------------------------------------------------- LN(gelu(u)) forward
// out = LayerNorm(act(x)) * gamma + beta, x bf16, out bf16.  NCH chunks of 8 columns per thread (n <= NCH * 2048).
template <int NCH>
__global__ void __launch_bounds__(ROW_THREADS)
act_layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ld_x, int act, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out, long long ld_out,
                         int rows, int n) {
    __shared__ float red[16];
    const float inv_n = 1.0f / static_cast<float>(n);
    uint4 raw[NCH];                                    // the next row's loads are in flight during this row's math
    auto fetch = [&](int row) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = (threadIdx.x + c * ROW_THREADS) * 8;
            raw[c] = (row < rows && col < n) ? *reinterpret_cast<const uint4*>(x + row * ld_x + col) : make_uint4(0, 0, 0, 0);
        }
    };
    fetch(blockIdx.x);
    // Packed f32x2 math throughout (FFMA2 / FMUL2 / FADD2): the scalar form spent ~13 FMA-pipe instructions per element, and that
    // pipe takes one warp instruction per 2 cycles and scheduler — ~97 us of the 130 us launch at n = 8192 against 83 us of HBM time.
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        uint64_t v[NCH][4];
        uint64_t sp = 0ull, qp = 0ull;
        const uint64_t kz = pack_f32x2(0.70710678118654752440f, 0.70710678118654752440f);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = (threadIdx.x + c * ROW_THREADS) * 8;
            const uint32_t w[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
            for (int u = 0; u < 4; ++u)        // bf16x2 -> f32x2: the low element is the word shifted up, the high one masked
                v[c][u] = pack_f32x2(__uint_as_float(w[u] << 16), __uint_as_float(w[u] & 0xffff0000u));
            if (col < n) {
                if (act == KX_ACT_GELU) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // a = x Phi(x), Phi(x) = 0.5 + copysign(0.5 - 0.5 erfc(|x| / sqrt2), x): the same expression the backward
                        // kernels differentiate (gelu_cdf's polynomial)
                        float xl, xh;
                        unpack_f32x2(v[c][u], xl, xh);
                        const uint64_t z = fmul2(pack_f32x2(fabsf(xl), fabsf(xh)), kz);
                        uint64_t t = ffma2(z, pack_f32x2(-0.003024620935320854f, -0.003024620935320854f),
                                           pack_f32x2(0.029882797971367836f, 0.029882797971367836f));
                        t = ffma2(t, z, pack_f32x2(-0.14901681244373322f, -0.14901681244373322f));
                        t = ffma2(t, z, pack_f32x2(-0.9183504581451416f, -0.9183504581451416f));
                        t = ffma2(t, z, pack_f32x2(-1.6279104948043823f, -1.6279104948043823f));
                        t = fmul2(t, z);
                        float t0, t1;
                        unpack_f32x2(t, t0, t1);
                        const uint64_t h = ffma2(pack_f32x2(ex2_approx(t0), ex2_approx(t1)), pack_f32x2(-0.5f, -0.5f), pack_f32x2(0.5f, 0.5f));
                        float h0, h1;
                        unpack_f32x2(h, h0, h1);
                        const uint64_t phi = fadd2(pack_f32x2(copysignf(h0, xl), copysignf(h1, xh)), pack_f32x2(0.5f, 0.5f));
                        v[c][u] = fmul2(v[c][u], phi);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { sp = fadd2(sp, v[c][u]); qp = ffma2(v[c][u], v[c][u], qp); }
            }
        }
        fetch(row + gridDim.x);
        float s, q, t0, t1;
        unpack_f32x2(sp, t0, t1); s = t0 + t1;
        unpack_f32x2(qp, t0, t1); q = t0 + t1;
        block_sum2(s, q, red);
        const float mean = s * inv_n;
        const float rstd = rsqrtf(fmaxf(q * inv_n - mean * mean, 0.f) + eps);
        const uint64_t rs2 = pack_f32x2(rstd, rstd), nmr2 = pack_f32x2(-mean * rstd, -mean * rstd);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = (threadIdx.x + c * ROW_THREADS) * 8;
            if (col < n) {
                const ulonglong2 g0 = *reinterpret_cast<const ulonglong2*>(gamma + col), g1 = *reinterpret_cast<const ulonglong2*>(gamma + col + 4);
                const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(beta + col), b1 = *reinterpret_cast<const ulonglong2*>(beta + col + 4);
                const uint64_t g[4] = {g0.x, g0.y, g1.x, g1.y}, bb[4] = {b0.x, b0.y, b1.x, b1.y};
                uint32_t ow[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float o0, o1;
                    unpack_f32x2(ffma2(ffma2(v[c][u], rs2, nmr2), g[u], bb[u]), o0, o1);
                    ow[u] = pack_bf16(o0, o1);
                }
                *reinterpret_cast<uint4*>(out + row * ld_out + col) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
        }
    }
}

// ----------------------------------------------------------------------------- LayerNorm backward
// y = LN(a) * gamma + beta with a = act(x) (+ pre_add).  Given dy (bf16), with g = dy * gamma:
//   da = rstd * (g - mean(g) - xhat * mean(g * xhat));  dx = da * act'(x)
//   d(gamma) += dy * xhat, d(beta) += dy  (per-CTA partials [grid][n], folded by colpartials_reduce_kernel)
// One pass, ONE block reduction per row: sum a, sum a^2, sum g, sum g*a give mean, rstd and
//   mean(g * xhat) = rstd * (sum(g*a) - mean * sum(g)) / n.
// Rows travel through a 3-stage shared-memory ring filled by 1-D bulk copies (cp.async.bulk + mbarrier, issued by
// thread 0 two rows ahead): ncu showed the register-prefetch version latency-bound (24 warps/SM, one row in flight per
// CTA, 3.3 TB/s) and the wide-row version issue-bound on its shared-memory accumulators.  Every thread owns 8 columns
// (THREADS = 256 for n <= 2048, 1024 for n <= 8192), so d(gamma) / d(beta) / column-sum accumulators are registers.
// DX_F32: dx is the fp32 residual-stream gradient: dx_out = dres + dx (dres may alias dx_out), an optional bf16 copy
// dxb of dx_out is written (the A operand of the next backward GEMMs) and column sums of dx_out are accumulated as a
// third partial (the bias gradient of the Linear whose output was added to the stream at this point).
constexpr int LNB_STAGES = 3;

// Four sums over the warp in 6 shuffles instead of 20: the first two butterfly steps halve the number of values a lane
// carries (a lane keeps the pair / the value its own half will own and sends the other), the last three reduce the one that
// is left.  On return lanes [8v, 8v + 8) hold the warp total of value v (a, b, c, d = 0..3) in `a`.  Fixed order: bit-reproducible.
__device__ __forceinline__ void warp_sum4_owner(float& a, float b, float c, float d) {
    const int lane = threadIdx.x & 31;
    const bool hi16 = lane & 16, hi8 = lane & 8;
    float k0 = hi16 ? c : a, k1 = hi16 ? d : b;
    k0 += __shfl_xor_sync(0xffffffffu, hi16 ? a : c, 16);
    k1 += __shfl_xor_sync(0xffffffffu, hi16 ? b : d, 16);
    float k = hi8 ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    a = k;
}

template <int THREADS>
__device__ __forceinline__ void block_sum4_t(float& a, float& b, float& c, float& d, float* red) {
    constexpr int W = THREADS / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    warp_sum4_owner(a, b, c, d);
    if ((lane & 7) == 0) red[(lane >> 3) * W + w] = a;          // lanes 0 / 8 / 16 / 24 own value 0 / 1 / 2 / 3
    __syncthreads();
    // every warp folds the 4 x W partials the same way (W <= 32: one partial per lane and value) and broadcasts the totals
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    if (lane < W) { p0 = red[lane]; p1 = red[W + lane]; p2 = red[2 * W + lane]; p3 = red[3 * W + lane]; }
    warp_sum4_owner(p0, p1, p2, p3);
    a = __shfl_sync(0xffffffffu, p0, 0);
    b = __shfl_sync(0xffffffffu, p0, 8);
    c = __shfl_sync(0xffffffffu, p0, 16);
    d = __shfl_sync(0xffffffffu, p0, 24);
    __syncthreads();
}

__device__ __forceinline__ void lnb_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename XT, bool DX_F32, int THREADS, bool GELU>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 1)
layernorm_bwd_kernel(const XT* __restrict__ x, long long ld_x, const float* __restrict__ pre_add,
                     const __nv_bfloat16* __restrict__ dy, long long ld_dy, const float* __restrict__ gamma, float eps,
                     const float* dres, long long ld_dres, void* dx_out, long long ld_dx, __nv_bfloat16* __restrict__ dxb,
                     long long ld_dxb, float* __restrict__ part_gamma, float* __restrict__ part_beta,
                     float* __restrict__ part_col, int rows, int n, const DropSpec drop) {
    extern __shared__ __align__(128) uint8_t lnb_smem[];
    __shared__ float red[4 * (THREADS / 32)];
    __shared__ __align__(8) uint64_t full[LNB_STAGES];
    const uint32_t x_bytes = static_cast<uint32_t>(n) * sizeof(XT), dy_bytes = static_cast<uint32_t>(n) * 2;
    const bool has_res = DX_F32 && dres != nullptr;
    const uint32_t res_bytes = has_res ? static_cast<uint32_t>(n) * 4 : 0;
    const uint32_t stage_bytes = x_bytes + dy_bytes + res_bytes;           // multiples of 16 (n % 8 == 0)
    const float inv_n = 1.0f / static_cast<float>(n);
    const int col = threadIdx.x * 8;
    const bool live = col < n;

    auto issue = [&](int row, int s) {                  // thread 0 only
        uint8_t* st = lnb_smem + static_cast<size_t>(s) * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        lnb_bulk_load(st, x + static_cast<long long>(row) * ld_x, x_bytes, &full[s]);
        lnb_bulk_load(st + x_bytes, dy + static_cast<long long>(row) * ld_dy, dy_bytes, &full[s]);
        if (has_res) lnb_bulk_load(st + x_bytes + dy_bytes, dres + static_cast<long long>(row) * ld_dres, res_bytes, &full[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < LNB_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        for (int s = 0; s < LNB_STAGES; ++s) {
            const int row = blockIdx.x + s * gridDim.x;
            if (row < rows) issue(row, s);
        }
    }
    __syncthreads();

    constexpr bool PRE_ADD = (THREADS == 256) && !GELU;          // x + media_pos_emb[i] only exists on narrow rows
    float gm[8], pa[PRE_ADD ? 8 : 1], acc_g[8], acc_b[8], acc_c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { gm[u] = 0.f; acc_g[u] = 0.f; acc_b[u] = 0.f; acc_c[u] = 0.f; }
    if constexpr (PRE_ADD) {
#pragma unroll
        for (int u = 0; u < 8; ++u) pa[u] = 0.f;
    }
    if (live) {
        if constexpr (!GELU) load8(gamma + col, gm);
        if constexpr (PRE_ADD) {
            if (pre_add != nullptr) load8(pre_add + col, pa);
        }
    }

    int it = 0;
    for (int row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
        const int s = it % LNB_STAGES;
        const uint8_t* st = lnb_smem + static_cast<size_t>(s) * stage_bytes;
        mbar_wait_lean(&full[s], (it / LNB_STAGES) & 1);
        if constexpr (GELU) {
            // ---- LN(gelu(u)) backward on packed f32x2 pairs (FFMA2 / FMUL2 / FADD2): the n = 8192 launch was issue-bound at
            // ~40 scalar instructions per element (ncu: 2.1 TB/s, 33 % of the copy bandwidth); the pair form issues ~half of them.
            // live across the block reduction: x, dy, Phi(x) (24 registers) — a = x * Phi(x) is recomputed and gamma re-read
            // from L1 after it, or the 64-register budget of a 1024-thread CTA spills
            uint64_t x2[4], d2[4], c2[4];
            {
                uint4 qx = make_uint4(0u, 0u, 0u, 0u), qd = make_uint4(0u, 0u, 0u, 0u);
                if (live) {
                    qx = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(st) + col);
                    qd = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(st + x_bytes) + col);
                }
                const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {       // bf16x2 -> f32x2: the low element is the word shifted up, the high one masked
                    x2[p] = pack_f32x2(__uint_as_float(wx[p] << 16), __uint_as_float(wx[p] & 0xffff0000u));
                    d2[p] = pack_f32x2(__uint_as_float(wd[p] << 16), __uint_as_float(wd[p] & 0xffff0000u));
                }
            }
            uint64_t s1p = 0ull, s2p = 0ull, sgp = 0ull, sgap = 0ull;
            const uint64_t kz = pack_f32x2(0.70710678118654752440f, 0.70710678118654752440f);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float xl, xh_;
                unpack_f32x2(x2[p], xl, xh_);
                const uint64_t z = fmul2(pack_f32x2(fabsf(xl), fabsf(xh_)), kz);
                uint64_t q = ffma2(z, pack_f32x2(-0.003024620935320854f, -0.003024620935320854f),
                                   pack_f32x2(0.029882797971367836f, 0.029882797971367836f));
                q = ffma2(q, z, pack_f32x2(-0.14901681244373322f, -0.14901681244373322f));
                q = ffma2(q, z, pack_f32x2(-0.9183504581451416f, -0.9183504581451416f));
                q = ffma2(q, z, pack_f32x2(-1.6279104948043823f, -1.6279104948043823f));
                q = fmul2(q, z);
                float q0, q1;
                unpack_f32x2(q, q0, q1);
                // Phi(x) = 0.5 + copysign(0.5 - 0.5 * erfc(|x| / sqrt2), x)
                const uint64_t h = ffma2(pack_f32x2(ex2_approx(q0), ex2_approx(q1)), pack_f32x2(-0.5f, -0.5f), pack_f32x2(0.5f, 0.5f));
                float h0, h1;
                unpack_f32x2(h, h0, h1);
                c2[p] = fadd2(pack_f32x2(copysignf(h0, xl), copysignf(h1, xh_)), pack_f32x2(0.5f, 0.5f));
                const uint64_t av = fmul2(x2[p], c2[p]);
                const float2 g2 = live ? __ldg(reinterpret_cast<const float2*>(gamma + col) + p) : make_float2(0.f, 0.f);
                const uint64_t gv = fmul2(d2[p], pack_f32x2(g2.x, g2.y));
                s1p = fadd2(s1p, av);
                s2p = ffma2(av, av, s2p);
                sgp = fadd2(sgp, gv);
                sgap = ffma2(gv, av, sgap);
            }
            float s1, s2, sg, sga, t0, t1;
            unpack_f32x2(s1p, t0, t1); s1 = t0 + t1;
            unpack_f32x2(s2p, t0, t1); s2 = t0 + t1;
            unpack_f32x2(sgp, t0, t1); sg = t0 + t1;
            unpack_f32x2(sgap, t0, t1); sga = t0 + t1;
            block_sum4_t<THREADS>(s1, s2, sg, sga, red);     // (its barriers also order this row's smem reads before the refill)
            if (threadIdx.x == 0) {
                const int nxt = row + LNB_STAGES * gridDim.x;
                if (nxt < rows) issue(nxt, s);
            }
            const float mean = s1 * inv_n;
            const float rstd = rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + eps);
            const float mg = sg * inv_n;
            const float mgx = rstd * (sga - mean * sg) * inv_n;       // mean(g * xhat)
            if (live) {
                const uint64_t rs2 = pack_f32x2(rstd, rstd), nmr2 = pack_f32x2(-mean * rstd, -mean * rstd);
                const uint64_t nmg2 = pack_f32x2(-mg, -mg), nmgx2 = pack_f32x2(-mgx, -mgx);
                uint32_t ow[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const uint64_t xh = ffma2(fmul2(x2[p], c2[p]), rs2, nmr2);                           // xhat
                    const float2 g2 = __ldg(reinterpret_cast<const float2*>(gamma + col) + p);
                    const uint64_t gv = fmul2(d2[p], pack_f32x2(g2.x, g2.y));
                    uint64_t o = fmul2(ffma2(xh, nmgx2, fadd2(gv, nmg2)), rs2);                          // LayerNorm backward
                    const uint64_t xx = fmul2(fmul2(x2[p], x2[p]), pack_f32x2(-0.72134752044448170368f, -0.72134752044448170368f));
                    float e0, e1;
                    unpack_f32x2(xx, e0, e1);
                    const uint64_t pdf = fmul2(pack_f32x2(ex2_approx(e0), ex2_approx(e1)), pack_f32x2(0.3989422804014327f, 0.3989422804014327f));
                    o = fmul2(o, ffma2(x2[p], pdf, c2[p]));                                              // * gelu'(x) = Phi + x phi
                    float o0, o1, v0, v1;
                    unpack_f32x2(o, o0, o1);
                    ow[p] = pack_bf16(o0, o1);
                    unpack_f32x2(ffma2(d2[p], xh, pack_f32x2(acc_g[2 * p], acc_g[2 * p + 1])), v0, v1);
                    acc_g[2 * p] = v0; acc_g[2 * p + 1] = v1;
                    unpack_f32x2(fadd2(d2[p], pack_f32x2(acc_b[2 * p], acc_b[2 * p + 1])), v0, v1);
                    acc_b[2 * p] = v0; acc_b[2 * p + 1] = v1;
                    // column sums of what the next GEMM reads: the bf16-rounded values
                    acc_c[2 * p] += __uint_as_float(ow[p] << 16);
                    acc_c[2 * p + 1] += __uint_as_float(ow[p] & 0xffff0000u);
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dx_out) + static_cast<long long>(row) * ld_dx + col) =
                    make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
            continue;
        }
        float xr[8], d[8], r[DX_F32 ? 8 : 1], a[8], cdf[GELU ? 8 : 1];
#pragma unroll
        for (int u = 0; u < 8; ++u) { xr[u] = 0.f; d[u] = 0.f; }
        if constexpr (DX_F32) {
#pragma unroll
            for (int u = 0; u < 8; ++u) r[u] = 0.f;
        }
        if (live) {
            load8(reinterpret_cast<const XT*>(st) + col, xr);
            load8(reinterpret_cast<const __nv_bfloat16*>(st + x_bytes) + col, d);
            if constexpr (DX_F32) {
                if (has_res) load8(reinterpret_cast<const float*>(st + x_bytes + dy_bytes) + col, r);
            }
        }
        float s1 = 0.f, s2 = 0.f, sg = 0.f, sga = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if constexpr (GELU) { cdf[u] = gelu_cdf(xr[u]); a[u] = xr[u] * cdf[u]; }
            else if constexpr (PRE_ADD) a[u] = xr[u] + pa[u];
            else a[u] = xr[u];
            const float gv = d[u] * gm[u];
            s1 += a[u]; s2 = fmaf(a[u], a[u], s2); sg += gv; sga = fmaf(gv, a[u], sga);
        }
        block_sum4_t<THREADS>(s1, s2, sg, sga, red);     // (its barriers also order this row's smem reads before the refill)
        if (threadIdx.x == 0) {
            const int nxt = row + LNB_STAGES * gridDim.x;
            if (nxt < rows) issue(nxt, s);
        }
        const float mean = s1 * inv_n;
        const float rstd = rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + eps);
        const float mg = sg * inv_n;
        const float mgx = rstd * (sga - mean * sg) * inv_n;       // mean(g * xhat)
        if (live) {
            float o[8], xh[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                xh[u] = (a[u] - mean) * rstd;
                o[u] = rstd * (d[u] * gm[u] - mg - xh[u] * mgx);
                if constexpr (GELU)                                // gelu'(x) = Phi(x) + x * phi(x)
                    o[u] *= fmaf(xr[u], 0.3989422804014327f * ex2_approx(-0.72134752044448170368f * xr[u] * xr[u]), cdf[u]);
                acc_g[u] = fmaf(d[u], xh[u], acc_g[u]);
                acc_b[u] += d[u];
            }
            if constexpr (DX_F32) {
                if (has_res)
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] += r[u];
                store8(reinterpret_cast<float*>(dx_out) + static_cast<long long>(row) * ld_dx + col, o);
                if (drop.thr != 0u) {
                    // the Linear output that was added to the stream here went through dropout in the forward (the GEMM
                    // epilogue's mask, regenerated): its gradient — the bf16 copy the backward GEMMs read and the bias
                    // column sums — is masked and rescaled; the stream gradient stored above is not
                    const uint32_t keep = drop_keep8(drop, static_cast<uint32_t>(row), static_cast<uint32_t>(col >> 3));
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] = ((keep >> u) & 1u) ? o[u] * drop.inv_keep : 0.f;
                }
                if (dxb != nullptr) store8(dxb + static_cast<long long>(row) * ld_dxb + col, o);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc_c[u] += o[u];
            } else {
                store8(reinterpret_cast<__nv_bfloat16*>(dx_out) + static_cast<long long>(row) * ld_dx + col, o);
                // column sums of what the next GEMM reads: the bf16-rounded values
#pragma unroll
                for (int u = 0; u < 8; ++u) acc_c[u] += __bfloat162float(__float2bfloat16_rn(o[u]));
            }
        }
    }
    if (live) {
        const long long o = static_cast<long long>(blockIdx.x) * n + col;
        store8(part_gamma + o, acc_g);
        store8(part_beta + o, acc_b);
        if (part_col != nullptr) store8(part_col + o, acc_c);
    }
}

// ----------------------------------------------------------------------------- LN(gelu(u)) backward, wide rows (2048 < n <= 8192)
// The ffn_layernorm backward of the decoder (n = 8192, 24 launches per step).  Same math as the GELU branch above, organised around
// what ncu showed of it (1024 threads x 8 columns: 44 instructions per element, 46 % issue-active, 92 bytes of spills at the
// 64-register budget, two block barriers per row with every warp in the same phase):
//   * 512 threads x 16 columns (two 16-byte chunks per thread): the block reduction costs half as much per element and the
//     128-register budget holds the three column accumulators AND gamma (no L1 re-reads);
//   * two rows in flight per CTA, software-pipelined: an iteration runs phase A of row k+1 (Phi, a = u Phi, the four sums ->
//     per-warp partials in smem) and phase B of row k (totals -> dx, column accumulators), ONE __syncthreads per row, and the
//     fold of row k's partials (LDS + 10 shuffles) is independent of phase A, so ptxas interleaves the two;
//   * nothing of a row is live in registers across the barrier: phase B re-reads u / dy from the ring slot and Phi from a
//     thread-private fp32 stash in shared memory (written by phase A; same thread, same addresses: no hazard).
// Ring: LNW_STAGES slots of (u, dy) filled by bulk copies two to three rows ahead; a slot is refilled after the barrier that
// ends the iteration of its phase B.
constexpr int LNW_THREADS = 512;
constexpr int LNW_STAGES = 5;
constexpr int LNW_STASH_BYTES = 4 * LNW_THREADS * 16;         // per buffer: [chunk][half][thread] float4

template <bool FULL>                                          // FULL: n == 16 * LNW_THREADS, every thread owns two live chunks
__global__ void __launch_bounds__(LNW_THREADS, 1)
ln_gelu_bwd_wide_kernel(const __nv_bfloat16* __restrict__ x, long long ld_x, const __nv_bfloat16* __restrict__ dy, long long ld_dy,
                        const float* __restrict__ gamma, float eps, __nv_bfloat16* __restrict__ dx, long long ld_dx,
                        float* __restrict__ part_gamma, float* __restrict__ part_beta, float* __restrict__ part_col, int rows, int n) {
    extern __shared__ __align__(128) uint8_t lnw_smem[];
    __shared__ float red[2][4 * (LNW_THREADS / 32)];
    __shared__ __align__(8) uint64_t full[LNW_STAGES];
    constexpr int W = LNW_THREADS / 32;
    const uint32_t row_bytes = static_cast<uint32_t>(n) * 2;  // one operand row (bf16); a slot holds u then dy
    const uint32_t stage_bytes = 2 * row_bytes;
    uint8_t* stash = lnw_smem + static_cast<size_t>(LNW_STAGES) * stage_bytes;
    const float inv_n = 1.0f / static_cast<float>(n);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col[2] = {tid * 8, (LNW_THREADS + tid) * 8};
    const bool live[2] = {FULL || col[0] < n, FULL || col[1] < n};
    const int grid = gridDim.x;
    const int n_rows = (rows - static_cast<int>(blockIdx.x) + grid - 1) / grid;      // rows blockIdx.x + k * grid, k < n_rows

    auto issue = [&](int k) {                           // thread 0 only
        const int s = k % LNW_STAGES;                   // (== the slot phase B of row k - LNW_STAGES just released)
        const long long row = blockIdx.x + static_cast<long long>(k) * grid;
        uint8_t* st = lnw_smem + static_cast<size_t>(s) * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        lnb_bulk_load(st, x + row * ld_x, row_bytes, &full[s]);
        lnb_bulk_load(st + row_bytes, dy + row * ld_dy, row_bytes, &full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < LNW_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        for (int k = 0; k < LNW_STAGES && k < n_rows; ++k) issue(k);
    }

    uint64_t gm2[2][4], acc_g[2][4], acc_b[2][4], acc_c[2][4];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        float g[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) g[u] = 0.f;
        if (live[c]) load8(gamma + col[c], g);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            gm2[c][p] = pack_f32x2(g[2 * p], g[2 * p + 1]);
            acc_g[c][p] = 0ull; acc_b[c][p] = 0ull; acc_c[c][p] = 0ull;
        }
    }
    __syncthreads();

    // bf16x2 word -> f32x2: the low element is the word shifted up, the high one masked
    auto widen = [](uint32_t w) { return pack_f32x2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); };

    // ---- phase A of row k: Phi(u) -> stash[k & 1], per-warp partial sums of (a, a^2, g, g a) -> red[k & 1]
    auto phase_a = [&](int k, int slot) {
        const uint8_t* st = lnw_smem + static_cast<size_t>(slot) * stage_bytes;
        uint8_t* sb = stash + (k & 1) * LNW_STASH_BYTES;
        uint64_t s1p = 0ull, s2p = 0ull, sgp = 0ull, sgap = 0ull;
        const uint64_t kz = pack_f32x2(0.70710678118654752440f, 0.70710678118654752440f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint4 qx = make_uint4(0u, 0u, 0u, 0u), qd = make_uint4(0u, 0u, 0u, 0u);
            if (live[c]) {
                qx = *reinterpret_cast<const uint4*>(st + col[c] * 2);
                qd = *reinterpret_cast<const uint4*>(st + row_bytes + col[c] * 2);
            }
            const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w};
            uint64_t phi[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const uint64_t x2 = widen(wx[p]), d2 = widen(wd[p]);
                float xl, xh_;
                unpack_f32x2(x2, xl, xh_);
                const uint64_t z = fmul2(pack_f32x2(fabsf(xl), fabsf(xh_)), kz);
                uint64_t q = ffma2(z, pack_f32x2(-0.003024620935320854f, -0.003024620935320854f),
                                   pack_f32x2(0.029882797971367836f, 0.029882797971367836f));
                q = ffma2(q, z, pack_f32x2(-0.14901681244373322f, -0.14901681244373322f));
                q = ffma2(q, z, pack_f32x2(-0.9183504581451416f, -0.9183504581451416f));
                q = ffma2(q, z, pack_f32x2(-1.6279104948043823f, -1.6279104948043823f));
                q = fmul2(q, z);
                float q0, q1;
                unpack_f32x2(q, q0, q1);
                // Phi(x) = 0.5 + copysign(0.5 - 0.5 * erfc(|x| / sqrt2), x)
                const uint64_t h = ffma2(pack_f32x2(ex2_approx(q0), ex2_approx(q1)), pack_f32x2(-0.5f, -0.5f), pack_f32x2(0.5f, 0.5f));
                float h0, h1;
                unpack_f32x2(h, h0, h1);
                phi[p] = fadd2(pack_f32x2(copysignf(h0, xl), copysignf(h1, xh_)), pack_f32x2(0.5f, 0.5f));
                const uint64_t av = fmul2(x2, phi[p]);
                const uint64_t gv = fmul2(d2, gm2[c][p]);
                s1p = fadd2(s1p, av);
                s2p = ffma2(av, av, s2p);
                sgp = fadd2(sgp, gv);
                sgap = ffma2(gv, av, sgap);
            }
            *reinterpret_cast<ulonglong2*>(sb + ((c * 2 + 0) * LNW_THREADS + tid) * 16) = make_ulonglong2(phi[0], phi[1]);
            *reinterpret_cast<ulonglong2*>(sb + ((c * 2 + 1) * LNW_THREADS + tid) * 16) = make_ulonglong2(phi[2], phi[3]);
        }
        float s1, s2, sg, sga, t0, t1;
        unpack_f32x2(s1p, t0, t1); s1 = t0 + t1;
        unpack_f32x2(s2p, t0, t1); s2 = t0 + t1;
        unpack_f32x2(sgp, t0, t1); sg = t0 + t1;
        unpack_f32x2(sgap, t0, t1); sga = t0 + t1;
        warp_sum4_owner(s1, s2, sg, sga);
        if ((lane & 7) == 0) red[k & 1][(lane >> 3) * W + warp] = s1;      // lanes 0 / 8 / 16 / 24 own value 0 / 1 / 2 / 3
    };

    // ---- row totals from the per-warp partials (every warp folds them the same way: bit-identical in all threads)
    struct Totals { float mean, rstd, mg, mgx; };
    auto fold = [&](int k) {
        const float* r = red[k & 1];
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (lane < W) { p0 = r[lane]; p1 = r[W + lane]; p2 = r[2 * W + lane]; p3 = r[3 * W + lane]; }
        warp_sum4_owner(p0, p1, p2, p3);
        const float s1 = __shfl_sync(0xffffffffu, p0, 0), s2 = __shfl_sync(0xffffffffu, p0, 8);
        const float sg = __shfl_sync(0xffffffffu, p0, 16), sga = __shfl_sync(0xffffffffu, p0, 24);
        Totals t;
        t.mean = s1 * inv_n;
        t.rstd = rsqrtf(fmaxf(s2 * inv_n - t.mean * t.mean, 0.f) + eps);
        t.mg = sg * inv_n;
        t.mgx = t.rstd * (sga - t.mean * sg) * inv_n;              // mean(g * xhat)
        return t;
    };

    // ---- phase B of row k: dx = LN backward * gelu'(u), column accumulators
    auto phase_b = [&](int k, int slot, const Totals& t) {
        const uint8_t* st = lnw_smem + static_cast<size_t>(slot) * stage_bytes;
        const uint8_t* sb = stash + (k & 1) * LNW_STASH_BYTES;
        const long long row = blockIdx.x + static_cast<long long>(k) * grid;
        const uint64_t rs2 = pack_f32x2(t.rstd, t.rstd), nmr2 = pack_f32x2(-t.mean * t.rstd, -t.mean * t.rstd);
        const uint64_t nmg2 = pack_f32x2(-t.mg, -t.mg), nmgx2 = pack_f32x2(-t.mgx, -t.mgx);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (!live[c]) continue;
            const uint4 qx = *reinterpret_cast<const uint4*>(st + col[c] * 2);
            const uint4 qd = *reinterpret_cast<const uint4*>(st + row_bytes + col[c] * 2);
            const ulonglong2 f0 = *reinterpret_cast<const ulonglong2*>(sb + ((c * 2 + 0) * LNW_THREADS + tid) * 16);
            const ulonglong2 f1 = *reinterpret_cast<const ulonglong2*>(sb + ((c * 2 + 1) * LNW_THREADS + tid) * 16);
            const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w};
            const uint64_t phi[4] = {f0.x, f0.y, f1.x, f1.y};
            uint32_t ow[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const uint64_t x2 = widen(wx[p]), d2 = widen(wd[p]);
                const uint64_t xh = ffma2(fmul2(x2, phi[p]), rs2, nmr2);                             // xhat
                const uint64_t gv = fmul2(d2, gm2[c][p]);
                uint64_t o = fmul2(ffma2(xh, nmgx2, fadd2(gv, nmg2)), rs2);                          // LayerNorm backward
                const uint64_t xx = fmul2(fmul2(x2, x2), pack_f32x2(-0.72134752044448170368f, -0.72134752044448170368f));
                float e0, e1;
                unpack_f32x2(xx, e0, e1);
                const uint64_t pdf = fmul2(pack_f32x2(ex2_approx(e0), ex2_approx(e1)), pack_f32x2(0.3989422804014327f, 0.3989422804014327f));
                o = fmul2(o, ffma2(x2, pdf, phi[p]));                                                // * gelu'(x) = Phi + x phi
                float o0, o1;
                unpack_f32x2(o, o0, o1);
                ow[p] = pack_bf16(o0, o1);
                acc_g[c][p] = ffma2(d2, xh, acc_g[c][p]);
                acc_b[c][p] = fadd2(d2, acc_b[c][p]);
                acc_c[c][p] = fadd2(widen(ow[p]), acc_c[c][p]);   // column sums of what the next GEMM reads: the bf16-rounded values
            }
            *reinterpret_cast<uint4*>(dx + row * ld_dx + col[c]) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
    };

    mbar_wait_lean(&full[0], 0);
    phase_a(0, 0);
    __syncthreads();
    int sb_slot = 0, sa_slot = 1 % LNW_STAGES;          // ring slots of row k (phase B) and row k + 1 (phase A)
    uint32_t sa_par = (1 / LNW_STAGES) & 1;
    for (int k = 0; k + 1 < n_rows; ++k) {
        mbar_wait_lean(&full[sa_slot], sa_par);
        const Totals t = fold(k);
        phase_a(k + 1, sa_slot);
        phase_b(k, sb_slot, t);
        __syncthreads();      // row k+1's partials complete; every read of row k's slot is done
        if (tid == 0 && k + LNW_STAGES < n_rows) issue(k + LNW_STAGES);
        sb_slot = sa_slot;
        if (++sa_slot == LNW_STAGES) { sa_slot = 0; sa_par ^= 1u; }
    }
    {
        const Totals t = fold(n_rows - 1);
        phase_b(n_rows - 1, sb_slot, t);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (!live[c]) continue;
        const long long o = static_cast<long long>(blockIdx.x) * n + col[c];
        *reinterpret_cast<ulonglong2*>(part_gamma + o) = make_ulonglong2(acc_g[c][0], acc_g[c][1]);
        *reinterpret_cast<ulonglong2*>(part_gamma + o + 4) = make_ulonglong2(acc_g[c][2], acc_g[c][3]);
        *reinterpret_cast<ulonglong2*>(part_beta + o) = make_ulonglong2(acc_b[c][0], acc_b[c][1]);
        *reinterpret_cast<ulonglong2*>(part_beta + o + 4) = make_ulonglong2(acc_b[c][2], acc_b[c][3]);
        if (part_col != nullptr) {
            *reinterpret_cast<ulonglong2*>(part_col + o) = make_ulonglong2(acc_c[c][0], acc_c[c][1]);
            *reinterpret_cast<ulonglong2*>(part_col + o + 4) = make_ulonglong2(acc_c[c][2], acc_c[c][3]);
        }
    }
}

// out_k[col] (+)= sum_p part_k[p][col] for up to three partial matrices (blockIdx.y = k).  A CTA owns 32 columns; its 8
// warps stride the partial rows (128-byte coalesced row segments, four loads in flight), then fold through smem.
__global__ void __launch_bounds__(256)
colpartials_reduce_kernel(const float* __restrict__ part, long long part_stride, int parts, int n, float* out0, float* out1,
                          float* out2, int accumulate) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const float* src = part + static_cast<long long>(blockIdx.y) * part_stride;
    float* out = blockIdx.y == 0 ? out0 : (blockIdx.y == 1 ? out1 : out2);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (col < n) {
        int p = w;
        for (; p + 24 < parts; p += 32) {
            s0 += src[static_cast<long long>(p) * n + col];
            s1 += src[static_cast<long long>(p + 8) * n + col];
            s2 += src[static_cast<long long>(p + 16) * n + col];
            s3 += src[static_cast<long long>(p + 24) * n + col];
        }
        for (; p < parts; p += 8) s0 += src[static_cast<long long>(p) * n + col];
    }
    red[w][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (w == 0 && col < n) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][lane];
        out[col] = accumulate ? out[col] + s : s;
    }
}

// ----------------------------------------------------------------------------- column sums (bias gradients)
// out[n] += sum_m x[m][n], x bf16.  Grid (ceil(N/256), row splits): a warp reads 256 consecutive columns of a row
// (16 bytes per lane), the 8 warps of a CTA take 8 rows at a time, four row groups in flight per thread.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows, int n, float* __restrict__ out) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + lane * 8;
    const int rows_per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    if (col < n) {                                       // n % 8 == 0: a lane's 8 columns are all inside or all outside
        int r = r0 + w;
        for (; r + 24 < r1; r += 32) {
            uint4 q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const uint4*>(x + static_cast<long long>(r + 8 * i) * ld + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q[i]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 f = __bfloat1622float2(h[u]);
                    acc[2 * u] += f.x; acc[2 * u + 1] += f.y;
                }
            }
        }
        for (; r < r1; r += 8) {
            float v[8];
            load8(x + static_cast<long long>(r) * ld + col, v);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] += v[u];
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) red[w][lane * 8 + u] = acc[u];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < n) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

// ----------------------------------------------------------------------------- xPos backward
// Forward (GEMM epilogue): o0 = x0*c - x1*s, o1 = x1*c + x0*s on column pairs of q (upscale tables) and k
// (downscale tables).  Backward: dx0 = d0*c + d1*s, dx1 = d1*c - d0*s, in place on the q|k column blocks.
__global__ void __launch_bounds__(256)
xpos_bwd_kernel(__nv_bfloat16* __restrict__ dqkv, long long ld, int rows, int d_model, int seq_len,
                const float* __restrict__ q_cos, const float* __restrict__ q_sin, const float* __restrict__ k_cos,
                const float* __restrict__ k_sin) {
    const int vec_per_row = (2 * d_model) >> 3;
    const long long total = static_cast<long long>(rows) * vec_per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / vec_per_row);
        const int col = static_cast<int>(i - static_cast<long long>(row) * vec_per_row) * 8;
        const int t = row % seq_len;
        const bool is_k = col >= d_model;
        const int j0 = (col & 63) >> 1;
        const float4 c = __ldg(reinterpret_cast<const float4*>((is_k ? k_cos : q_cos) + t * 32 + j0));
        const float4 s = __ldg(reinterpret_cast<const float4*>((is_k ? k_sin : q_sin) + t * 32 + j0));
        __nv_bfloat16* p = dqkv + static_cast<long long>(row) * ld + col;
        float v[8];
        load8(p, v);
        const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
        float o[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            o[2 * u] = v[2 * u] * cc[u] + v[2 * u + 1] * ss[u];
            o[2 * u + 1] = v[2 * u + 1] * cc[u] - v[2 * u] * ss[u];
        }
        store8(p, o);
    }
}

// ----------------------------------------------------------------------------- cross-entropy over text rows
struct SpliceRowsT { int count; int start[KX_MAX_IMAGES]; };

// Next-token targets (see kx_loss_targets in the header).  One thread per row of the spliced sequence.
__device__ __forceinline__ bool is_marker(int ti, int n_img, const SpliceRowsT& img) {
    for (int i = 0; i < img.count; ++i) {
        const int p = img.start[i] - i * n_img;              // image i sits in front of text token p
        if (ti == p - 1 || ti == p) return true;             // `<image>`, `</image>` (model.py:70-77)
    }
    return false;
}

__global__ void __launch_bounds__(256)
loss_targets_kernel(const long long* __restrict__ tokens, int batch, int t_text, int n_img, const SpliceRowsT img, int rule,
                    long long ignore_token, long long* __restrict__ targets, float* __restrict__ count) {
    const int T = t_text + n_img * img.count;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    int valid = 0;
    if (row < batch * T) {
        const int b = row / T, t = row - b * T;
        int ti = t;
        bool is_img = false, next_is_img = false;
        for (int i = 0; i < img.count; ++i) {
            if (t >= img.start[i]) {
                if (t < img.start[i] + n_img) is_img = true;
                ti -= n_img;
            }
            if (t + 1 == img.start[i]) next_is_img = true;
        }
        long long tgt = -100;
        if (!is_img) {
            if (rule == KX_LOSS_NEXT_TOKEN) {
                if (!next_is_img && ti + 1 < t_text) tgt = tokens[static_cast<long long>(b) * t_text + ti + 1];
            } else if (!is_marker(ti, n_img, img)) {
                int nx = ti + 1;
                while (nx < t_text && is_marker(nx, n_img, img)) ++nx;
                if (nx < t_text) tgt = tokens[static_cast<long long>(b) * t_text + nx];
            }
        }
        if (tgt >= 0 && tgt == ignore_token) tgt = -100;
        targets[row] = tgt;
        valid = tgt >= 0 ? 1 : 0;
    }
    if (count != nullptr) {
        const unsigned mask = __ballot_sync(0xffffffffu, valid);
        if ((threadIdx.x & 31) == 0 && mask) atomicAdd(count, static_cast<float>(__popc(mask)));
    }
}

// One CTA per row.  Pass 1: online (max, sum exp) over the fp32 logits; pass 2 (L2-resident re-read): d(logits) =
// (softmax - onehot) / max(*count, 1) as bf16, zero for rows without a target and for the pad columns [vocab, ld_d).
// loss_acc[0] += sum of row losses, loss_acc[1] += number of rows with a target.
__global__ void __launch_bounds__(256)
ce_fwd_bwd_kernel(const float* __restrict__ logits, long long ld_l, const long long* __restrict__ targets, int vocab,
                  const float* __restrict__ count, __nv_bfloat16* __restrict__ dlogits,
                  long long ld_d, float* __restrict__ loss_acc, int* __restrict__ err_flag) {
    __shared__ float red[16];
    const int row = blockIdx.x;
    long long tgt = targets[row];
    if (tgt >= vocab) {
        if (err_flag != nullptr && threadIdx.x == 0) atomicExch(err_flag, 1);
        tgt = -1;
    }
    const float inv_count = 1.0f / fmaxf(count != nullptr ? __ldg(count) : 1.0f, 1.0f);
    __nv_bfloat16* drow = dlogits != nullptr ? dlogits + static_cast<long long>(row) * ld_d : nullptr;
    if (tgt < 0) {
        if (drow != nullptr)
            for (int c = threadIdx.x * 8; c < ld_d; c += 256 * 8) *reinterpret_cast<uint4*>(drow + c) = make_uint4(0, 0, 0, 0);
        return;
    }
    const float* lrow = logits + static_cast<long long>(row) * ld_l;
    float m = -INFINITY, s = 0.f;
    for (int c = threadIdx.x * 2; c < vocab; c += 512) {          // rows are 8-byte aligned (ld even): float2 loads
        const float2 v = *reinterpret_cast<const float2*>(lrow + c);
        const float hi = (c + 1 < vocab) ? v.y : -INFINITY;
        const float nm = fmaxf(m, fmaxf(v.x, hi));
        s = s * __expf(m - nm) + __expf(v.x - nm) + __expf(hi - nm);
        m = nm;
    }
    // combine (m, s) pairs: first the max, then the rescaled sums
    float gm = m, dummy = 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gm;
    __syncthreads();
    gm = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) gm = fmaxf(gm, red[i]);
    __syncthreads();
    float ss = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
    block_sum2(ss, dummy, red);
    const float lse = gm + logf(ss);
    if (threadIdx.x == 0) {
        atomicAdd(loss_acc, lse - lrow[tgt]);
        atomicAdd(loss_acc + 1, 1.0f);
    }
    if (drow == nullptr) return;
    for (int c = threadIdx.x * 8; c < ld_d; c += 256 * 8) {
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
            float2 v = make_float2(0.f, 0.f);
            if (c + u < vocab) v = *reinterpret_cast<const float2*>(lrow + c + u);
            o[u] = (c + u < vocab) ? (__expf(v.x - lse) - ((c + u) == tgt ? 1.f : 0.f)) * inv_count : 0.f;
            o[u + 1] = (c + u + 1 < vocab) ? (__expf(v.y - lse) - ((c + u + 1) == tgt ? 1.f : 0.f)) * inv_count : 0.f;
        }
        store8(drow + c, o);
    }
}

// ----------------------------------------------------------------------------- embedding backward
// d(embed)[token] += dx0[row], d(pos)[t + 2] += dx0[row] for every row (image rows take part in the position sum only;
// their dx0 is also the gradient of the image_proj output).  The padding row (index padding_idx) gets no gradient,
// as nn.Embedding(padding_idx=...).  fp32 vector atomics.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const float* __restrict__ dx0, const long long* __restrict__ tokens, int t_text, int n_img,
                 const SpliceRowsT img, int dim, int vocab, int padding_idx, int alias, float* __restrict__ d_embed,
                 float* __restrict__ d_pos) {
    const int T = t_text + n_img * img.count;
    const int b = blockIdx.x / T, t = blockIdx.x - b * T;
    int ti = t;
    bool is_img = false;
    for (int i = 0; i < img.count; ++i) {
        if (t >= img.start[i]) {
            if (t < img.start[i] + n_img) is_img = true;
            ti -= n_img;
        }
    }
    const float4* src = reinterpret_cast<const float4*>(dx0 + static_cast<long long>(blockIdx.x) * dim);
    float4* pos = d_pos != nullptr ? reinterpret_cast<float4*>(d_pos + static_cast<long long>(t + 2) * dim) : nullptr;
    // alias_positions (kx_embed_splice_pos): the row of text token ti also received pos[ti + 2]
    float4* pos1 = (alias && !is_img && d_pos != nullptr) ? reinterpret_cast<float4*>(d_pos + static_cast<long long>(ti + 2) * dim) : nullptr;
    float4* emb = nullptr;
    if (!is_img && d_embed != nullptr) {
        const long long tok = tokens[static_cast<long long>(b) * t_text + ti];
        if (tok >= 0 && tok < vocab && tok != padding_idx) emb = reinterpret_cast<float4*>(d_embed + tok * dim);
    }
    for (int i = threadIdx.x; i < dim / 4; i += blockDim.x) {
        const float4 v = src[i];
        if (pos != nullptr) atomicAdd(pos + i, v);
        if (pos1 != nullptr) atomicAdd(pos1 + i, v);
        if (emb != nullptr) atomicAdd(emb + i, v);
    }
}

// ----------------------------------------------------------------------------- dropout on an fp32 matrix (decoder input)
// x[r, c] = keep(r, c) ? x[r, c] / keep_prob : 0, in place: torchscale's `x = dropout(x)` at the end of forward_embedding
// (the call at model.py:242-244 whose [0] is the decoder input) and, applied to the gradient, its backward.
__global__ void __launch_bounds__(256)
dropout_f32_kernel(float* __restrict__ x, long long ld, int rows, int cols, const DropSpec drop) {
    const int vec = cols >> 3;
    const long long total = static_cast<long long>(rows) * vec;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int r = static_cast<int>(i / vec), c8 = static_cast<int>(i - static_cast<long long>(r) * vec);
        float v[8];
        float* p = x + static_cast<long long>(r) * ld + c8 * 8;
        load8(p, v);
        const uint32_t keep = drop_keep8(drop, static_cast<uint32_t>(r), static_cast<uint32_t>(c8));
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ((keep >> u) & 1u) ? v[u] * drop.inv_keep : 0.f;
        store8(p, v);
    }
}

// ----------------------------------------------------------------------------- gradient norm + optimizers
// Deterministic: block b writes its partial to scratch[b]; sumsq_finish_kernel folds the partials in index order.  (An
// atomicAdd per block would make the gradient norm — hence the clip coefficient, hence every parameter — depend on block
// scheduling: data-parallel replicas holding bit-identical gradients would drift apart by ulps per step.)
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ scratch) {
    __shared__ float red[16];
    float s = 0.f, dummy = 0.f;
    const long long nv = n >> 2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = nv << 2; i < n; ++i) s = fmaf(g[i], g[i], s);
    block_sum2(s, dummy, red);
    if (threadIdx.x == 0) scratch[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
sumsq_finish_kernel(const float* __restrict__ scratch, int parts, float* __restrict__ out) {
    __shared__ float red[16];
    float s = 0.f, dummy = 0.f;
    for (int i = threadIdx.x; i < parts; i += 256) s += scratch[i];          // fixed assignment, fixed order
    block_sum2(s, dummy, red);
    if (threadIdx.x == 0) *out += s;
}

// scale = min(1, max_norm / (sqrt(sumsq * pre_scale^2) + 1e-6)) * pre_scale   (torch.nn.utils.clip_grad_norm_;
// pre_scale = 1/world averages all-reduced gradient sums); norm_out = the unclipped norm
__global__ void clip_scale_kernel(const float* __restrict__ sumsq, float max_norm, float pre_scale, float* __restrict__ scale,
                                  float* __restrict__ norm_out) {
    const float norm = sqrtf(*sumsq) * pre_scale;
    if (norm_out != nullptr) *norm_out = norm;
    float c = 1.0f;
    if (max_norm > 0.f) c = fminf(1.0f, max_norm / (norm + 1e-6f));
    *scale = c * pre_scale;
}

// torch.optim.AdamW (decoupled weight decay), fp32 master weights and moments; writes the bf16 operand copy.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ wb, long long n, float lr, float b1, float b2, float eps, float wd, float bc1,
             float bc2, const float* __restrict__ gscale) {
    const float gs = gscale != nullptr ? *gscale : 1.0f;
    const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * gs;
        float pi = p[i];
        pi *= 1.0f - lr * wd;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        pi -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
        p[i] = pi;
        if (wb != nullptr) wb[i] = __float2bfloat16_rn(pi);
    }
}

// lion_pytorch.Lion (train.py:375-379): p *= 1 - lr*wd; p -= lr * sign(b1*m + (1-b1)*g); m = b2*m + (1-b2)*g
__global__ void __launch_bounds__(256)
lion_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, __nv_bfloat16* __restrict__ wb,
            long long n, float lr, float b1, float b2, float wd, const float* __restrict__ gscale) {
    const float gs = gscale != nullptr ? *gscale : 1.0f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * gs, mi = m[i];
        float pi = p[i] * (1.0f - lr * wd);
        const float u = b1 * mi + (1.0f - b1) * gi;
        pi -= lr * ((u > 0.f) ? 1.0f : (u < 0.f ? -1.0f : 0.0f));
        m[i] = b2 * mi + (1.0f - b2) * gi;
        p[i] = pi;
        if (wb != nullptr) wb[i] = __float2bfloat16_rn(pi);
    }
}

// ----------------------------------------------------------------------------- perceiver feed-forward GELU (no LayerNorm behind it)
// mid = gelu(u);  du = dmid * gelu'(u)   (bf16, 8 elements per thread).  QUICK: CLIP's x * sigmoid(1.702 x) (the last ViT
// layer when it is fine-tuned) instead of the erf form.
__device__ __forceinline__ float quick_gelu_grad(float x) {       // s + 1.702 x s (1 - s),  s = sigmoid(1.702 x)
    const float s = 1.0f / (1.0f + __expf(-1.702f * x));
    return s * fmaf(1.702f * x, 1.0f - s, 1.0f);
}
template <bool QUICK>
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ out, long long nvec) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float v[8];
        load8(u + i * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = QUICK ? quick_gelu(v[k]) : gelu_exact(v[k]);
        store8(out + i * 8, v);
    }
}
template <bool QUICK>
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ dmid, __nv_bfloat16* __restrict__ du, long long nvec) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float v[8], d[8];
        load8(u + i * 8, v);
        load8(dmid + i * 8, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] *= QUICK ? quick_gelu_grad(v[k]) : gelu_grad(v[k]);
        store8(du + i * 8, d);
    }
}

// dst[r] (+)= src[(r / grp_rows) * grp_stride + grp_off + r % grp_rows]  (bf16 dst; src fp32 or bf16): picks the image rows
// out of the decoder-input gradient, and the media / latent rows out of the [media | latents] gradient of the resampler.
template <typename ST>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const ST* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst, long long ld_dst, int rows, int n,
                   int grp_rows, int grp_stride, int grp_off, int accumulate) {
    const int vec_per_row = n >> 3;
    const long long total = static_cast<long long>(rows) * vec_per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / vec_per_row), col = static_cast<int>(i - static_cast<long long>(r) * vec_per_row) * 8;
        const long long sr = grp_rows > 0 ? static_cast<long long>(r / grp_rows) * grp_stride + grp_off + r % grp_rows : r;
        float v[8];
        load8(src + sr * ld_src + col, v);
        if (accumulate) {
            float o[8];
            load8(dst + static_cast<long long>(r) * ld_dst + col, o);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += o[k];
        }
        store8(dst + static_cast<long long>(r) * ld_dst + col, v);
    }
}

// out[c] (+)= sum_r src[r][c]  (fp32; gradient of the broadcast perceiver latents: sum over the images)
__global__ void __launch_bounds__(256)
sum_rows_f32_kernel(const float* __restrict__ src, long long ld, int rows, long long n, float* __restrict__ out, int accumulate) {
    const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (c >= n) return;
    float s = accumulate ? out[c] : 0.f;
    for (int r = 0; r < rows; ++r) s += src[static_cast<long long>(r) * ld + c];
    out[c] = s;
}

static int grid_for(long long n, int sms, int per_thread = 1) {
    const long long blocks = (n / per_thread + 255) / 256;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(blocks, static_cast<long long>(sms) * 8)));
}

static bool fill_splice(SpliceRowsT& img, const int* host_img_rows, int img_count, int n_img, int T, const char* who) {
    if (img_count < 0 || img_count > KX_MAX_IMAGES || (img_count > 0 && !host_img_rows)) {
        set_error("%s: bad img_count %d", who, img_count);
        return false;
    }
    img.count = img_count;
    for (int i = 0; i < img_count; ++i) {
        const int s = host_img_rows[i];
        if (s < 0 || s + n_img > T || (i > 0 && s < host_img_rows[i - 1] + n_img)) {
            set_error("%s: image %d cannot start at spliced row %d (T=%d, %d rows per image)", who, i, s, T, n_img);
            return false;
        }
        img.start[i] = s;
    }
    return true;
}

}  // namespace kx

using namespace kx;

#define KX_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int kx_ln_bwd_partials(int rows) {
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    return std::max(1, std::min(rows, sms * 3));
}

extern "C" int kx_act_layernorm_fwd(const void* x_bf16, long long ld_x, int act, const float* gamma, const float* beta,
                                    float eps, void* out_bf16, long long ld_out, int rows, int n, cudaStream_t stream) {
    if (!x_bf16 || !gamma || !beta || !out_bf16 || rows <= 0 || n <= 0 || (n % 8) || n > 8192 || (ld_x % 8) || (ld_out % 8) ||
        !KX_ALIGNED16(x_bf16) || !KX_ALIGNED16(out_bf16) || !KX_ALIGNED16(gamma) || !KX_ALIGNED16(beta) ||
        (act != KX_ACT_NONE && act != KX_ACT_GELU)) {
        set_error("kx_act_layernorm_fwd: bad argument (rows=%d n=%d: n %% 8 == 0, n <= 8192, 16-byte aligned rows)", rows, n);
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const int grid = std::min(rows, sms * 8);
    auto xp = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
    auto op = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    if (n <= 2048) act_layernorm_fwd_kernel<1><<<grid, ROW_THREADS, 0, stream>>>(xp, ld_x, act, gamma, beta, eps, op, ld_out, rows, n);
    else if (n <= 4096) act_layernorm_fwd_kernel<2><<<grid, ROW_THREADS, 0, stream>>>(xp, ld_x, act, gamma, beta, eps, op, ld_out, rows, n);
    else act_layernorm_fwd_kernel<4><<<grid, ROW_THREADS, 0, stream>>>(xp, ld_x, act, gamma, beta, eps, op, ld_out, rows, n);
    return check_launch("kx_act_layernorm_fwd");
}

extern "C" int kx_layernorm_bwd(const void* x, int x_is_bf16, long long ld_x, const float* pre_add, int act, const void* dy_bf16, long long ld_dy,
                                const float* gamma, float eps, const float* dres, long long ld_dres, void* dx,
                                int dx_is_f32, long long ld_dx, void* dxb_bf16, long long ld_dxb, float* partials,
                                int n_partials, float* d_gamma, float* d_beta, float* d_colsum, int accumulate, int rows,
                                int n, float drop_p, unsigned int drop_site, unsigned long long drop_seed, cudaStream_t stream) {
    if (drop_p != 0.f && (!(drop_p > 0.f && drop_p < 1.f) || !dx_is_f32)) {
        set_error("kx_layernorm_bwd: dropout needs 0 < p < 1 and the fp32 residual form (dx_is_f32)");
        return KX_ERR_ARG;
    }
    const DropSpec drop = make_drop_spec(drop_p, drop_site, drop_seed);
    if (!x || !dy_bf16 || !gamma || !dx || !partials || !d_gamma || !d_beta || rows <= 0 || n <= 0 || (n % 8) || n > 8192 ||
        (ld_x % 8) || (ld_dy % 8) || (ld_dx % 8) || !KX_ALIGNED16(x) || !KX_ALIGNED16(dy_bf16) || !KX_ALIGNED16(dx) ||
        !KX_ALIGNED16(gamma) || !KX_ALIGNED16(partials) || (dres && (!KX_ALIGNED16(dres) || (ld_dres % 4) || !dx_is_f32)) ||
        (dxb_bf16 && (!KX_ALIGNED16(dxb_bf16) || (ld_dxb % 8) || !dx_is_f32)) || (act != KX_ACT_NONE && act != KX_ACT_GELU) ||
        (pre_add && (n > 2048 || act != KX_ACT_NONE || !KX_ALIGNED16(pre_add)))) {
        set_error("kx_layernorm_bwd: bad argument (rows=%d n=%d; n %% 8 == 0, n <= 8192, 16-byte aligned rows; dres / dxb need fp32 dx)", rows, n);
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const bool wide = n > 2048;
    const int grid = std::max(1, std::min(rows, sms * (wide ? 1 : 3)));
    if (n_partials < grid) { set_error("kx_layernorm_bwd: partials buffer holds %d rows, needs kx_ln_bwd_partials(rows) = %d", n_partials, grid); return KX_ERR_ARG; }
    if ((ld_x * (x_is_bf16 ? 2 : 4)) % 16 || (ld_dy * 2) % 16 || (dres && (ld_dres * 4) % 16)) {
        set_error("kx_layernorm_bwd: rows must be 16-byte aligned for the bulk copies");
        return KX_ERR_ARG;
    }
    float* pg = partials;
    float* pb = partials + static_cast<long long>(grid) * n;
    float* pc = d_colsum ? partials + 2ll * grid * n : nullptr;
    auto dyp = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
    auto dxbp = reinterpret_cast<__nv_bfloat16*>(dxb_bf16);
    const int smem = LNB_STAGES * n * ((x_is_bf16 ? 2 : 4) + 2 + ((dx_is_f32 && dres) ? 4 : 0));
#define KX_LNB(XT, F32, TH, GE)                                                                                           \
    {                                                                                                                     \
        auto kern = layernorm_bwd_kernel<XT, F32, TH, GE>;                                                                \
        static int attr = 0;                                                                                              \
        if (smem > attr) {                                                                                                \
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {           \
                set_error("kx_layernorm_bwd: cudaFuncSetAttribute(smem=%d) failed", smem);                                \
                return KX_ERR_LAUNCH;                                                                                     \
            }                                                                                                             \
            attr = smem;                                                                                                  \
        }                                                                                                                 \
        kern<<<grid, TH, smem, stream>>>(reinterpret_cast<const XT*>(x), ld_x, pre_add, dyp, ld_dy, gamma, eps, dres,     \
                                         ld_dres, dx, ld_dx, dxbp, ld_dxb, pg, pb, pc, rows, n, drop);                    \
    }
#define KX_LNB_N(XT, F32, GE)                                                                                             \
    { if (!wide) KX_LNB(XT, F32, 256, GE) else KX_LNB(XT, F32, 1024, GE) }
    if (act == KX_ACT_GELU) {
        if (!x_is_bf16 || dx_is_f32) { set_error("kx_layernorm_bwd: the GELU form takes a bf16 pre-activation and writes a bf16 gradient"); return KX_ERR_ARG; }
        if (wide) {
            // ffn_layernorm of the decoder: the two-rows-in-flight kernel (512 threads x 16 columns)
            const int wsmem = LNW_STAGES * n * 4 + 2 * LNW_STASH_BYTES;
            static std::once_flag once;
            static bool ok = false;
            std::call_once(once, [] {
                constexpr int most = LNW_STAGES * 8192 * 4 + 2 * LNW_STASH_BYTES;
                ok = cudaFuncSetAttribute(ln_gelu_bwd_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most) == cudaSuccess &&
                     cudaFuncSetAttribute(ln_gelu_bwd_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most) == cudaSuccess;
            });
            if (!ok) { set_error("kx_layernorm_bwd: cudaFuncSetAttribute failed (wide GELU form)"); return KX_ERR_LAUNCH; }
            auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
            auto dxp = reinterpret_cast<__nv_bfloat16*>(dx);
            if (n == 16 * LNW_THREADS)
                ln_gelu_bwd_wide_kernel<true><<<grid, LNW_THREADS, wsmem, stream>>>(xb, ld_x, dyp, ld_dy, gamma, eps, dxp, ld_dx, pg, pb, pc, rows, n);
            else
                ln_gelu_bwd_wide_kernel<false><<<grid, LNW_THREADS, wsmem, stream>>>(xb, ld_x, dyp, ld_dy, gamma, eps, dxp, ld_dx, pg, pb, pc, rows, n);
        } else KX_LNB(__nv_bfloat16, false, 256, true)
    } else if (x_is_bf16) { if (dx_is_f32) KX_LNB_N(__nv_bfloat16, true, false) else KX_LNB_N(__nv_bfloat16, false, false) }
    else { if (dx_is_f32) KX_LNB_N(float, true, false) else KX_LNB_N(float, false, false) }
#undef KX_LNB_N
#undef KX_LNB
    int st = check_launch("kx_layernorm_bwd");
    if (st != KX_OK) return st;
    colpartials_reduce_kernel<<<dim3((n + 31) / 32, pc ? 3 : 2), 256, 0, stream>>>(pg, static_cast<long long>(grid) * n, grid, n,
                                                                                   d_gamma, d_beta, d_colsum, accumulate);
    return check_launch("kx_layernorm_bwd (partials)");
}

extern "C" int kx_colsum_bf16(const void* x_bf16, long long ld, int rows, int n, float* out, cudaStream_t stream) {
    if (!x_bf16 || !out || rows <= 0 || n <= 0 || (n % 8) || (ld % 8) || !KX_ALIGNED16(x_bf16)) {
        set_error("kx_colsum_bf16: bad argument (rows=%d n=%d; n %% 8 == 0 and 16-byte aligned rows)", rows, n);
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const int cb = (n + 255) / 256;
    const int splits = std::max(1, std::min((rows + 63) / 64, (sms * 8 + cb - 1) / cb));
    colsum_bf16_kernel<<<dim3(cb, splits), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x_bf16), ld, rows, n, out);
    return check_launch("kx_colsum_bf16");
}

extern "C" int kx_xpos_bwd(void* dqkv_bf16, long long ld, int rows, int d_model, int seq_len, const float* q_cos,
                           const float* q_sin, const float* k_cos, const float* k_sin, cudaStream_t stream) {
    if (!dqkv_bf16 || !q_cos || !q_sin || !k_cos || !k_sin || rows <= 0 || d_model <= 0 || (d_model % 64) || seq_len <= 0 ||
        (ld % 8) || !KX_ALIGNED16(dqkv_bf16)) {
        set_error("kx_xpos_bwd: bad argument");
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const long long total = static_cast<long long>(rows) * (2 * d_model / 8);
    xpos_bwd_kernel<<<grid_for(total, sms), 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), ld, rows, d_model,
                                                              seq_len, q_cos, q_sin, k_cos, k_sin);
    return check_launch("kx_xpos_bwd");
}

extern "C" int kx_loss_targets(const long long* tokens, int batch, int t_text, const int* host_img_rows, int img_count, int n_img,
                               int rule, long long ignore_token, long long* targets, float* count, cudaStream_t stream) {
    if (!tokens || !targets || batch <= 0 || t_text <= 0 || n_img < 0 || (rule != KX_LOSS_REFERENCE && rule != KX_LOSS_NEXT_TOKEN)) {
        set_error("kx_loss_targets: bad argument (batch=%d t_text=%d n_img=%d rule=%d)", batch, t_text, n_img, rule);
        return KX_ERR_ARG;
    }
    const int T = t_text + n_img * img_count;
    SpliceRowsT img = {};
    if (!fill_splice(img, host_img_rows, img_count, n_img, T, "kx_loss_targets")) return KX_ERR_ARG;
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    loss_targets_kernel<<<(batch * T + 255) / 256, 256, 0, stream>>>(tokens, batch, t_text, n_img, img, rule, ignore_token, targets, count);
    return check_launch("kx_loss_targets");
}

extern "C" int kx_ce_fwd_bwd(const float* logits, long long ld_logits, const long long* targets, int rows, int vocab,
                             const float* count, void* dlogits_bf16, long long ld_dlogits, float* loss_acc, int* err_flag,
                             cudaStream_t stream) {
    if (!logits || !targets || !loss_acc || rows <= 0 || vocab <= 0 || (ld_logits % 2) || (reinterpret_cast<uintptr_t>(logits) & 7) ||
        (dlogits_bf16 && ((ld_dlogits % 8) || ld_dlogits < vocab || !KX_ALIGNED16(dlogits_bf16)))) {
        set_error("kx_ce_fwd_bwd: bad argument (logits rows 8-byte aligned; dlogits rows 16-byte aligned, ld >= vocab)");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    ce_fwd_bwd_kernel<<<rows, 256, 0, stream>>>(logits, ld_logits, targets, vocab, count,
                                                reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), ld_dlogits, loss_acc, err_flag);
    return check_launch("kx_ce_fwd_bwd");
}

extern "C" int kx_embed_bwd(const float* dx0, const long long* tokens, int batch, int t_text, const int* host_img_rows,
                            int img_count, int n_img, int dim, int vocab, int padding_idx, int alias_positions, float* d_embed,
                            float* d_pos, cudaStream_t stream) {
    if (!dx0 || !tokens || batch <= 0 || t_text <= 0 || (dim % 4) || !KX_ALIGNED16(dx0) || (d_embed && !KX_ALIGNED16(d_embed)) ||
        (d_pos && !KX_ALIGNED16(d_pos))) {
        set_error("kx_embed_bwd: bad argument");
        return KX_ERR_ARG;
    }
    const int T = t_text + n_img * img_count;
    SpliceRowsT img = {};
    if (!fill_splice(img, host_img_rows, img_count, n_img, T, "kx_embed_bwd")) return KX_ERR_ARG;
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    embed_bwd_kernel<<<batch * T, 256, 0, stream>>>(dx0, tokens, t_text, n_img, img, dim, vocab, padding_idx, alias_positions ? 1 : 0,
                                                    d_embed, d_pos);
    return check_launch("kx_embed_bwd");
}

extern "C" int kx_dropout_f32(float* x, long long ld, int rows, int cols, float drop_p, unsigned int drop_site,
                              unsigned long long drop_seed, cudaStream_t stream) {
    if (!x || rows <= 0 || cols <= 0 || (cols % 8) || (ld % 4) || !KX_ALIGNED16(x) || !(drop_p > 0.f && drop_p < 1.f)) {
        set_error("kx_dropout_f32: bad argument (cols %% 8 == 0, 16-byte aligned rows, 0 < p < 1)");
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    dropout_f32_kernel<<<grid_for(static_cast<long long>(rows) * (cols / 8), sms), 256, 0, stream>>>(x, ld, rows, cols,
                                                                                                   make_drop_spec(drop_p, drop_site, drop_seed));
    return check_launch("kx_dropout_f32");
}

extern "C" int kx_sumsq(const float* g, long long n, float* out, float* scratch, cudaStream_t stream) {
    if (!g || !out || !scratch || n <= 0 || !KX_ALIGNED16(g)) { set_error("kx_sumsq: bad argument"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const int parts = std::min(grid_for(n, sms, 4), KX_SUMSQ_SCRATCH);
    sumsq_kernel<<<parts, 256, 0, stream>>>(g, n, scratch);
    int st = check_launch("kx_sumsq");
    if (st != KX_OK) return st;
    sumsq_finish_kernel<<<1, 256, 0, stream>>>(scratch, parts, out);
    return check_launch("kx_sumsq");
}

extern "C" int kx_clip_scale(const float* sumsq, float max_norm, float pre_scale, float* scale_out, float* norm_out,
                             cudaStream_t stream) {
    if (!sumsq || !scale_out) { set_error("kx_clip_scale: null pointer"); return KX_ERR_ARG; }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    clip_scale_kernel<<<1, 1, 0, stream>>>(sumsq, max_norm, pre_scale, scale_out, norm_out);
    return check_launch("kx_clip_scale");
}

extern "C" int kx_adamw_step(float* p, const float* g, float* m, float* v, void* w_bf16, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, const float* grad_scale,
                             cudaStream_t stream) {
    if (!p || !g || !m || !v || n <= 0 || step < 1) { set_error("kx_adamw_step: bad argument"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step)), bc2 = 1.0f - powf(beta2, static_cast<float>(step));
    adamw_kernel<<<grid_for(n, sms), 256, 0, stream>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(w_bf16), n, lr, beta1, beta2,
                                                       eps, weight_decay, bc1, bc2, grad_scale);
    return check_launch("kx_adamw_step");
}

extern "C" int kx_lion_step(float* p, const float* g, float* m, void* w_bf16, long long n, float lr, float beta1, float beta2,
                            float weight_decay, const float* grad_scale, cudaStream_t stream) {
    if (!p || !g || !m || n <= 0) { set_error("kx_lion_step: bad argument"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    lion_kernel<<<grid_for(n, sms), 256, 0, stream>>>(p, g, m, reinterpret_cast<__nv_bfloat16*>(w_bf16), n, lr, beta1, beta2,
                                                      weight_decay, grad_scale);
    return check_launch("kx_lion_step");
}

extern "C" int kx_act_fwd(const void* u_bf16, void* out_bf16, long long n, int act, cudaStream_t stream) {
    if (!u_bf16 || !out_bf16 || n <= 0 || (n % 8) || !KX_ALIGNED16(u_bf16) || !KX_ALIGNED16(out_bf16) ||
        (act != KX_ACT_GELU && act != KX_ACT_QUICK_GELU)) {
        set_error("kx_act_fwd: bad argument");
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const auto* u = reinterpret_cast<const __nv_bfloat16*>(u_bf16);
    auto* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    if (act == KX_ACT_QUICK_GELU) gelu_fwd_kernel<true><<<grid_for(n / 8, sms), 256, 0, stream>>>(u, out, n / 8);
    else gelu_fwd_kernel<false><<<grid_for(n / 8, sms), 256, 0, stream>>>(u, out, n / 8);
    return check_launch("kx_act_fwd");
}

extern "C" int kx_act_bwd(const void* u_bf16, const void* dmid_bf16, void* du_bf16, long long n, int act, cudaStream_t stream) {
    if (!u_bf16 || !dmid_bf16 || !du_bf16 || n <= 0 || (n % 8) || !KX_ALIGNED16(u_bf16) || !KX_ALIGNED16(dmid_bf16) || !KX_ALIGNED16(du_bf16) ||
        (act != KX_ACT_GELU && act != KX_ACT_QUICK_GELU)) {
        set_error("kx_act_bwd: bad argument");
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const auto* u = reinterpret_cast<const __nv_bfloat16*>(u_bf16);
    const auto* dm = reinterpret_cast<const __nv_bfloat16*>(dmid_bf16);
    auto* du = reinterpret_cast<__nv_bfloat16*>(du_bf16);
    if (act == KX_ACT_QUICK_GELU) gelu_bwd_kernel<true><<<grid_for(n / 8, sms), 256, 0, stream>>>(u, dm, du, n / 8);
    else gelu_bwd_kernel<false><<<grid_for(n / 8, sms), 256, 0, stream>>>(u, dm, du, n / 8);
    return check_launch("kx_act_bwd");
}

extern "C" int kx_gelu_fwd(const void* u_bf16, void* out_bf16, long long n, cudaStream_t stream) {
    return kx_act_fwd(u_bf16, out_bf16, n, KX_ACT_GELU, stream);
}

extern "C" int kx_gelu_bwd(const void* u_bf16, const void* dmid_bf16, void* du_bf16, long long n, cudaStream_t stream) {
    return kx_act_bwd(u_bf16, dmid_bf16, du_bf16, n, KX_ACT_GELU, stream);
}

extern "C" int kx_gather_rows(const void* src, int src_is_f32, long long ld_src, void* dst_bf16, long long ld_dst, int rows, int n,
                              int grp_rows, int grp_stride, int grp_off, int accumulate, cudaStream_t stream) {
    if (!src || !dst_bf16 || rows <= 0 || n <= 0 || (n % 8) || (ld_src % 8) || (ld_dst % 8) || !KX_ALIGNED16(src) || !KX_ALIGNED16(dst_bf16) ||
        grp_rows < 0) {
        set_error("kx_gather_rows: bad argument (n %% 8 == 0, 16-byte aligned rows)");
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const long long total = static_cast<long long>(rows) * (n / 8);
    auto d = reinterpret_cast<__nv_bfloat16*>(dst_bf16);
    if (src_is_f32) gather_rows_kernel<float><<<grid_for(total, sms), 256, 0, stream>>>(reinterpret_cast<const float*>(src), ld_src, d, ld_dst, rows, n, grp_rows, grp_stride, grp_off, accumulate);
    else gather_rows_kernel<__nv_bfloat16><<<grid_for(total, sms), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), ld_src, d, ld_dst, rows, n, grp_rows, grp_stride, grp_off, accumulate);
    return check_launch("kx_gather_rows");
}

extern "C" int kx_sum_rows_f32(const float* src, long long ld, int rows, long long n, float* out, int accumulate, cudaStream_t stream) {
    if (!src || !out || rows <= 0 || n <= 0) { set_error("kx_sum_rows_f32: bad argument"); return KX_ERR_ARG; }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    sum_rows_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, ld, rows, n, out, accumulate);
    return check_launch("kx_sum_rows_f32");
}
