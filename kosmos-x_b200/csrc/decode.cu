// Incremental decoding: one new token per sequence against a KV cache — torchscale's
// `incremental_state` path of Decoder / MultiheadAttention.  Every kernel here is HBM-bound: a decoding step streams
// the 2.55 GB of decoder weights and the KV cache once and does almost no arithmetic per byte, so the design rules are
// coalesced 16-byte loads, enough loads in flight per SM (Little's law: ~45 KB), grids larger than the SM count, and
// no host round trip between steps (the position lives in device memory, so one captured graph serves every step).
//
//   decode_linear_kernel     y[B,N] = epilogue(a[B,K] . W[N,K]^T), B <= 32.  W rows are the M dimension of
//                            mma.sync.m16n8k16 (the batch is the 8-wide N dimension), fragments are loaded straight
//                            from global memory in a k-permuted layout so that every thread issues full 16-byte loads;
//                            the LayerNorm in front of the Linear is folded exactly as in the prefill GEMM
//                            (rstd*(acc - mean*c) + d) with the row statistics accumulated from the fragments the
//                            thread loads anyway.  tcgen05 needs a 128-row tile; at 8 rows the legacy warp MMA is the
//                            right tool and the kernel is bound by the weight stream either way.
//   decode_attention_kernel  flash-decoding: (batch, head, 128-key chunk) CTAs, partial (max, sum, o[64]) per chunk,
//                            the last CTA of a (batch, head) merges the chunks.
//   kv_cache_store_kernel    prompt pass: rotated k and v of a layer's q|k|v matrix -> the cache.
//   decode_embed_kernel      embedding row + learned position for the new token.
//   argmax_advance_kernel    greedy choice (or a forced token), history, position += 1.
#include "kx_internal.h"
#include "ptx.cuh"

namespace kx {

namespace {

struct DecLin {
    const __nv_bfloat16* a; long long lda; int B;
    const __nv_bfloat16* w; long long ldw; int N, K;
    const float* ln_c; const float* bias; float eps; int ln;
    int mode, act;
    void* out; long long ld_out; int out_f32; unsigned long long* argmax_keys;
    float* x; long long ld_x; __nv_bfloat16* xb; long long ld_xb;
    __nv_bfloat16* q_out; long long ld_q;
    __nv_bfloat16* k_cache; __nv_bfloat16* v_cache; int t_max; int d_model;
    const int* pos;
    const float *xq_cos, *xq_sin, *xk_cos, *xk_sin;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void bf16x2_stats(uint32_t v, float& s1, float& s2) {
    const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
    s1 += lo + hi;
    s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
}

// activations were written by the previous kernel of the chain (possibly still resident under PDL): plain coherent loads
__device__ __forceinline__ uint4 ld_act(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// weights are read exactly once per step: streaming loads, do not keep them in L1
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// Programmatic dependent launch: the next kernel of a decoding step is scheduled while this one drains; everything a
// kernel does before pdl_wait() may only touch memory no earlier kernel of the step writes (weights, tables).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

// A bulk-copy (cp.async.bulk -> smem) variant of decode_linear_kernel was measured in round 1 (1.059 ms per step against
// 1.028 ms: the step is bound by the per-kernel dependency chain, not by the in-kernel stream rate) and removed in round 2.
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int check_chain(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("%s launch failed: %s", what, cudaGetErrorString(e));
        return KX_ERR_LAUNCH;
    }
    count_launch();
    return KX_OK;
}

}  // namespace

// One CTA = 16 rows of W (16 output features) x the whole K, 8 warps splitting K; NB groups of 8 batch rows.
// U = 32-wide k steps (2 x 16-byte weight loads each) a thread keeps in flight; MINB = CTAs per SM the registers allow.
// Wide N (>= one CTA per SM slot): U = 4, 3 CTAs/SM.  Narrow N (fewer tiles than SMs): one CTA per SM has to keep the
// whole 45 KB/SM in flight alone, U = 8.
template <int NB, int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
decode_linear_kernel(const DecLin p) {
    constexpr int NBC = NB * 8;
    __shared__ float red[8][16][NBC];
    __shared__ float st[8][NBC][2];
    __shared__ float fin[16][NBC];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * 16;
    const int steps = p.K >> 5;
    const int spw = (steps + 7) >> 3;
    const int s_begin = warp * spw, s_end = min(steps, s_begin + spw);

    const int r0 = min(n0 + g, p.N - 1), r1 = min(n0 + g + 8, p.N - 1);
    const uint4* w0 = reinterpret_cast<const uint4*>(p.w + static_cast<long long>(r0) * p.ldw) + t;
    const uint4* w1 = reinterpret_cast<const uint4*>(p.w + static_cast<long long>(r1) * p.ldw) + t;
    const uint4* ap[NB];
    bool aok[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const int b = nb * 8 + g;
        aok[nb] = b < p.B;
        ap[nb] = reinterpret_cast<const uint4*>(p.a + static_cast<long long>(aok[nb] ? b : 0) * p.lda) + t;
    }

    float acc[NB][4];
    float s1[NB], s2[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
        s1[nb] = s2[nb] = 0.f;
    }

    pdl_launch_dependents();
    // Epilogue operands of this thread's first output (batch row tid / 16, feature tid % 16) are requested up front so
    // that their DRAM round trips overlap the weight stream: the folded-LayerNorm / bias vectors are immutable (before
    // the dependency wait), the xPos table entries and the residual value right after it.
    const int e_n = min(n0 + (tid & 15), p.N - 1);
    const int e_b = tid >> 4;
    float e_lnc = 0.f, e_bias = 0.f, e_c = 1.f, e_s = 0.f, e_x = 0.f;
    if (tid < 16 * NBC) {
        if (p.ln) e_lnc = __ldg(p.ln_c + e_n);
        if (p.bias != nullptr) e_bias = __ldg(p.bias + e_n);
    }
    auto after_wait = [&]() {
        if (tid >= 16 * NBC || e_b >= p.B) return;
        if (p.mode == KX_DEC_QKV) {
            const int which = e_n / p.d_model;
            if (which < 2) {
                const int pos = *p.pos;
                const int j = (e_n & 63) >> 1;
                e_c = __ldg((which == 0 ? p.xq_cos : p.xk_cos) + pos * 32 + j);
                e_s = __ldg((which == 0 ? p.xq_sin : p.xk_sin) + pos * 32 + j);
            }
        } else if (p.mode == KX_DEC_RESIDUAL) {
            e_x = p.x[static_cast<long long>(e_b) * p.ld_x + e_n];
        }
    };
    bool waited = false;
    for (int s = s_begin; s < s_end; s += U) {
        uint4 wa[U], wb[U], av[U][NB];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = s + u < s_end;
            const int o = (s + u) * 4;                     // uint4 index of this step's first k (32 k = 4 x 16 bytes)
            wa[u] = ok ? ld_stream(w0 + o) : make_uint4(0, 0, 0, 0);
            wb[u] = ok ? ld_stream(w1 + o) : make_uint4(0, 0, 0, 0);
        }
        if (!waited) { pdl_wait(); waited = true; after_wait(); }   // weights do not depend on the previous kernel; `a` does
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = s + u < s_end;
            const int o = (s + u) * 4;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
                av[u][nb] = (ok && aok[nb]) ? ld_act(ap[nb] + o) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                // k permutation: the thread's 8 consecutive k are fed as logical (2t,2t+1 | 2t+8,2t+9) of two MMAs,
                // identically for the W (A operand) and activation (B operand) fragments
                mma_bf16_16816(acc[nb], wa[u].x, wb[u].x, wa[u].y, wb[u].y, av[u][nb].x, av[u][nb].y);
                mma_bf16_16816(acc[nb], wa[u].z, wb[u].z, wa[u].w, wb[u].w, av[u][nb].z, av[u][nb].w);
                if (p.ln) {
                    bf16x2_stats(av[u][nb].x, s1[nb], s2[nb]); bf16x2_stats(av[u][nb].y, s1[nb], s2[nb]);
                    bf16x2_stats(av[u][nb].z, s1[nb], s2[nb]); bf16x2_stats(av[u][nb].w, s1[nb], s2[nb]);
                }
            }
        }
    }

    if (!waited) { pdl_wait(); after_wait(); }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        red[warp][g][nb * 8 + 2 * t] = acc[nb][0];
        red[warp][g][nb * 8 + 2 * t + 1] = acc[nb][1];
        red[warp][g + 8][nb * 8 + 2 * t] = acc[nb][2];
        red[warp][g + 8][nb * 8 + 2 * t + 1] = acc[nb][3];
        if (p.ln) {
            float a1 = s1[nb], a2 = s2[nb];
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
            if (t == 0) { st[warp][nb * 8 + g][0] = a1; st[warp][nb * 8 + g][1] = a2; }
        }
    }
    __syncthreads();

    // cross-warp sum (fixed order) + folded LayerNorm + bias; thread o -> (batch b = o / 16, feature r = o % 16)
    for (int o = tid; o < 16 * NBC; o += 256) {
        const int b = o >> 4, r = o & 15;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][r][b];
        const int n = min(n0 + r, p.N - 1);
        if (p.ln) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { a1 += st[w][b][0]; a2 += st[w][b][1]; }
            const float inv_n = 1.0f / static_cast<float>(p.K);
            const float mean = a1 * inv_n;
            const float var = fmaxf(a2 * inv_n - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.eps);
            v = fmaf(-mean * rstd, o == tid ? e_lnc : __ldg(p.ln_c + n), v * rstd);
        }
        if (p.bias != nullptr) v += (o == tid ? e_bias : __ldg(p.bias + n));
        fin[r][b] = v;
    }
    __syncthreads();

    for (int o = tid; o < 16 * NBC; o += 256) {
        const int b = o >> 4, r = o & 15;
        const int n = n0 + r;
        const bool live = b < p.B && n < p.N;
        float v = fin[r][b];
        if (p.argmax_keys != nullptr) {
            // greedy choice fused into output_projection: order-preserving (value, lowest index wins) 64-bit keys,
            // reduced over the tile's 16 features (one half-warp per batch row), one atomicMax per (tile, row)
            unsigned long long key = 0ull;
            if (live) {
                const uint32_t u = __float_as_uint(v);
                key = (static_cast<unsigned long long>((u & 0x80000000u) ? ~u : (u | 0x80000000u)) << 32) |
                      static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(n));
            }
#pragma unroll
            for (int sft = 8; sft > 0; sft >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, sft);
                key = other > key ? other : key;
            }
            if (r == 0 && b < p.B) atomicMax(p.argmax_keys + b, key);
        }
        if (!live) continue;
        if (p.mode == KX_DEC_QKV) {
            const int which = n / p.d_model;
            const int col = n - which * p.d_model;
            const int pos = *p.pos;
            if (which < 2) {                               // xPos rotation of the (2j, 2j+1) pair at position pos
                const int j = (n & 63) >> 1;
                const float c = o == tid ? e_c : __ldg((which == 0 ? p.xq_cos : p.xk_cos) + pos * 32 + j);
                const float s = o == tid ? e_s : __ldg((which == 0 ? p.xq_sin : p.xk_sin) + pos * 32 + j);
                const float x0 = fin[r & ~1][b], x1 = fin[r | 1][b];
                v = (r & 1) ? fmaf(x1, c, x0 * s) : fmaf(x0, c, -(x1 * s));
            }
            if (which == 0) {
                p.q_out[static_cast<long long>(b) * p.ld_q + col] = __float2bfloat16_rn(v);
            } else if (pos < p.t_max) {
                __nv_bfloat16* dst = (which == 1 ? p.k_cache : p.v_cache);
                // cache layout [batch, heads, t_max, 64]: one (batch, head) is one contiguous stream for kx_decode_attn
                const int hh = col >> 6;
                dst[((static_cast<long long>(b) * (p.d_model >> 6) + hh) * p.t_max + pos) * 64 + (col & 63)] = __float2bfloat16_rn(v);
            }
        } else if (p.mode == KX_DEC_RESIDUAL) {
            float* px = p.x + static_cast<long long>(b) * p.ld_x + n;
            v += (o == tid ? e_x : *px);
            *px = v;
            p.xb[static_cast<long long>(b) * p.ld_xb + n] = __float2bfloat16_rn(v);
        } else {
            if (p.act == KX_ACT_GELU) v = gelu_erf(v);
            else if (p.act == KX_ACT_QUICK_GELU) v = quick_gelu(v);
            if (p.out_f32) reinterpret_cast<float*>(p.out)[static_cast<long long>(b) * p.ld_out + n] = v;
            else reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(b) * p.ld_out + n] = __float2bfloat16_rn(v);
        }
    }
}

// ----------------------------------------------------------------------------- decode attention
// grid (chunks, heads, batch), 128 threads.  A warp owns 32 keys of the chunk; 8 lanes share one key row (64 bf16 =
// 8 x 16 bytes), 4 keys per load instruction, all 8 K loads and all 8 V loads of a thread are issued up front.
constexpr int DA_CHUNK = 128;

__global__ void __launch_bounds__(128)
decode_attention_kernel(const __nv_bfloat16* __restrict__ q, long long ld_q, const __nv_bfloat16* __restrict__ k_cache,
                        const __nv_bfloat16* __restrict__ v_cache, int t_max, int d_model, const int* __restrict__ pos_ptr,
                        float scale_log2, float* __restrict__ part, int* __restrict__ counters,
                        __nv_bfloat16* __restrict__ out, long long ld_out, int max_chunks) {
    pdl_launch_dependents();
    pdl_wait();
    const int n_keys = min(*pos_ptr + 1, t_max);             // the new token attends to itself and everything before
    const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z, H = gridDim.y;
    const int n_act = (n_keys + DA_CHUNK - 1) / DA_CHUNK;
    if (chunk >= n_act) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & 7, kq = lane >> 3;
    __shared__ float sm_o[4][64];
    __shared__ float sm_ml[4][2];
    __shared__ int is_last;

    float qf[8];
    {
        const uint4 raw = ld_act(reinterpret_cast<const uint4*>(q + static_cast<long long>(b) * ld_q + h * 64) + sub);
        const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            qf[2 * i] = __uint_as_float(r[i] << 16) * scale_log2;
            qf[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u) * scale_log2;
        }
    }
    const int key0 = chunk * DA_CHUNK + warp * 32 + kq;
    const long long base = (static_cast<long long>(b) * H + h) * t_max * 64;
    uint4 kr[8], vr[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int j = key0 + it * 4;
        kr[it] = j < n_keys ? ld_act(reinterpret_cast<const uint4*>(k_cache + base + static_cast<long long>(j) * 64) + sub)
                            : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int j = key0 + it * 4;
        vr[it] = j < n_keys ? ld_act(reinterpret_cast<const uint4*>(v_cache + base + static_cast<long long>(j) * 64) + sub)
                            : make_uint4(0, 0, 0, 0);
    }
    float sc[8];
    float m = -INFINITY;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const uint32_t r[4] = {kr[it].x, kr[it].y, kr[it].z, kr[it].w};
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d = fmaf(qf[2 * i], __uint_as_float(r[i] << 16), d);
            d = fmaf(qf[2 * i + 1], __uint_as_float(r[i] & 0xffff0000u), d);
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        sc[it] = (key0 + it * 4 < n_keys) ? d : -INFINITY;
        m = fmaxf(m, sc[it]);
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    float l = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (m > -INFINITY) {                                    // warp-uniform: at least one valid key in this warp's slice
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const float pj = ex2_approx(sc[it] - m);       // exp2(-inf) = 0 for the masked tail
            l += pj;
            const uint32_t r[4] = {vr[it].x, vr[it].y, vr[it].z, vr[it].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o[2 * i] = fmaf(pj, __uint_as_float(r[i] << 16), o[2 * i]);
                o[2 * i + 1] = fmaf(pj, __uint_as_float(r[i] & 0xffff0000u), o[2 * i + 1]);
            }
        }
    }
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    }
    if (kq == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm_o[warp][sub * 8 + i] = o[i];
        if (sub == 0) { sm_ml[warp][0] = m; sm_ml[warp][1] = l; }
    }
    __syncthreads();
    // CTA partial -> scratch [b][h][chunk][66]
    float* my = part + ((static_cast<long long>(b) * H + h) * max_chunks + chunk) * 66;
    if (threadIdx.x < 64) {
        float M = fmaxf(fmaxf(sm_ml[0][0], sm_ml[1][0]), fmaxf(sm_ml[2][0], sm_ml[3][0]));
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float f = sm_ml[w][0] > -INFINITY ? ex2_approx(sm_ml[w][0] - M) : 0.f;
            L = fmaf(sm_ml[w][1], f, L);
            O = fmaf(sm_o[w][threadIdx.x], f, O);
        }
        my[threadIdx.x] = O;
        if (threadIdx.x == 0) { my[64] = M; my[65] = L; }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(counters + b * H + h, 1);
        is_last = (done == n_act - 1);
        if (is_last) counters[b * H + h] = 0;               // ready for the next step (stream order)
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 64) {
        const float* all = part + (static_cast<long long>(b) * H + h) * max_chunks * 66;
        float M = -INFINITY, L = 0.f, O = 0.f;
        for (int c0 = 0; c0 < n_act; c0 += 8) {             // 24 independent L2 loads in flight per pass
            float mc[8], lc[8], oc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool ok = c0 + i < n_act;
                const float* pc = all + (ok ? c0 + i : c0) * 66;
                mc[i] = ok ? __ldcg(pc + 64) : -INFINITY;
                lc[i] = __ldcg(pc + 65);
                oc[i] = __ldcg(pc + threadIdx.x);
            }
            float Mn = M;
#pragma unroll
            for (int i = 0; i < 8; ++i) Mn = fmaxf(Mn, mc[i]);
            const float f0 = M > -INFINITY ? ex2_approx(M - Mn) : 0.f;
            L *= f0; O *= f0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float f = mc[i] > -INFINITY ? ex2_approx(mc[i] - Mn) : 0.f;
                L = fmaf(lc[i], f, L);
                O = fmaf(oc[i], f, O);
            }
            M = Mn;
        }
        out[static_cast<long long>(b) * ld_out + h * 64 + threadIdx.x] = __float2bfloat16_rn(O / L);
    }
}

// ----------------------------------------------------------------------------- prompt pass: fill the cache
__global__ void __launch_bounds__(256)
kv_cache_store_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int batch, int seq_len, int d_model,
                      __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache, int t_max) {
    const int vec_per_row = d_model >> 3;
    const long long total = static_cast<long long>(batch) * seq_len * vec_per_row * 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = static_cast<int>(i % vec_per_row);
        const long long rw = i / vec_per_row;
        const int which = static_cast<int>(rw & 1);
        const long long row = rw >> 1;                     // b * seq_len + t
        const int b = static_cast<int>(row / seq_len), t = static_cast<int>(row - static_cast<long long>(b) * seq_len);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(qkv + row * ld + static_cast<long long>(which + 1) * d_model) + c);
        __nv_bfloat16* dst = which ? v_cache : k_cache;
        const int hh = c >> 3;                              // 8 x 16 bytes per 64-wide head
        reinterpret_cast<uint4*>(dst + ((static_cast<long long>(b) * (d_model >> 6) + hh) * t_max + t) * 64)[c & 7] = v;
    }
}

// ----------------------------------------------------------------------------- new-token embedding
__global__ void __launch_bounds__(256)
decode_embed_kernel(const long long* __restrict__ tok, const float* __restrict__ embed, int vocab,
                    const float* __restrict__ pos_tab, int pos_rows, const int* __restrict__ pos_ptr, int text_off, int dim,
                    float* __restrict__ x, __nv_bfloat16* __restrict__ xb, int* __restrict__ err_flag) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x;
    long long id = tok[b];
    if (id < 0 || id >= vocab) {
        if (threadIdx.x == 0 && err_flag != nullptr) atomicOr(err_flag, 1);
        id = 0;
    }
    int pr = *pos_ptr + 2;                            // "positions start from 2" 
    if (pr >= pos_rows) {
        if (threadIdx.x == 0 && err_flag != nullptr) atomicOr(err_flag, 2);
        pr = pos_rows - 1;
    }
    const float4* e = reinterpret_cast<const float4*>(embed + id * dim);
    const float4* pp = reinterpret_cast<const float4*>(pos_tab + static_cast<long long>(pr) * dim);
    // text_off >= 0: the sequence was embedded with alias_positions (kx_embed_splice_pos) — the text-index position first
    const float4* p1 = text_off >= 0 ? reinterpret_cast<const float4*>(pos_tab + static_cast<long long>(max(pr - text_off, 0)) * dim) : nullptr;
    for (int i = threadIdx.x; i < (dim >> 2); i += blockDim.x) {
        float4 a = __ldg(e + i);
        const float4 c = __ldg(pp + i);
        if (p1 != nullptr) { const float4 c1 = __ldg(p1 + i); a.x += c1.x; a.y += c1.y; a.z += c1.z; a.w += c1.w; }
        const float4 r = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
        reinterpret_cast<float4*>(x + static_cast<long long>(b) * dim)[i] = r;
        uint2 pk;
        pk.x = pack_bf16(r.x, r.y); pk.y = pack_bf16(r.z, r.w);
        reinterpret_cast<uint2*>(xb + static_cast<long long>(b) * dim)[i] = pk;
    }
}

// ----------------------------------------------------------------------------- greedy choice + bookkeeping
__global__ void __launch_bounds__(256)
argmax_advance_kernel(const float* __restrict__ logits, long long ld, int vocab, const long long* __restrict__ forced,
                      long long* __restrict__ tok_out, long long* __restrict__ history, int hist_ld,
                      int* __restrict__ pos_ptr, int* __restrict__ step_ptr, int* __restrict__ counter,
                      unsigned long long* __restrict__ keys, int batch) {
    pdl_launch_dependents();
    pdl_wait();
    if (keys != nullptr) {                                  // the LM head already reduced (value, index) keys: one CTA
        const int step = *step_ptr;
        __syncthreads();
        for (int b = threadIdx.x; b < batch; b += blockDim.x) {
            long long choice = static_cast<long long>(0xffffffffu - static_cast<uint32_t>(keys[b] & 0xffffffffull));
            keys[b] = 0ull;
            if (forced != nullptr && step < hist_ld) choice = forced[static_cast<long long>(b) * hist_ld + step];
            tok_out[b] = choice;
            if (history != nullptr && step < hist_ld) history[static_cast<long long>(b) * hist_ld + step] = choice;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            *step_ptr = step + 1;
            if (pos_ptr != nullptr) *pos_ptr += 1;
        }
        return;
    }
    const int b = blockIdx.x;
    const int step = *step_ptr;
    __shared__ float sv[8];
    __shared__ int si[8];
    const float* row = logits + static_cast<long long>(b) * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        const float v = row[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
        long long choice = bi;
        if (forced != nullptr && step < hist_ld) choice = forced[static_cast<long long>(b) * hist_ld + step];
        tok_out[b] = choice;
        if (history != nullptr && step < hist_ld) history[static_cast<long long>(b) * hist_ld + step] = choice;
        __threadfence();
        const int done = atomicAdd(counter, 1);
        if (done == static_cast<int>(gridDim.x) - 1) {      // every row has read `step`: advance
            *counter = 0;
            *step_ptr = step + 1;
            if (pos_ptr != nullptr) *pos_ptr += 1;
        }
    }
}

}  // namespace kx

using namespace kx;

extern "C" int kx_decode_linear(const void* a, long long lda, int batch, const void* w, long long ldw, int N, int K,
                                const kx_decode_linear_args* g, cudaStream_t stream) {
    if (!a || !w || !g || batch <= 0 || batch > KX_DECODE_MAX_BATCH || N <= 0 || K <= 0 || (K & 31) || (lda & 7) || (ldw & 7) ||
        lda < K || ldw < K || ((uintptr_t)a & 15) || ((uintptr_t)w & 15)) {
        set_error("kx_decode_linear: need 1 <= batch <= %d, K %% 32 == 0, 16-byte aligned rows (batch %d, N %d, K %d)",
                  KX_DECODE_MAX_BATCH, batch, N, K);
        return KX_ERR_ARG;
    }
    DecLin p{};
    p.a = reinterpret_cast<const __nv_bfloat16*>(a); p.lda = lda; p.B = batch;
    p.w = reinterpret_cast<const __nv_bfloat16*>(w); p.ldw = ldw; p.N = N; p.K = K;
    p.ln = g->ln_c != nullptr; p.ln_c = g->ln_c; p.bias = g->bias; p.eps = g->ln_eps;
    p.mode = g->mode; p.act = g->act;
    if (g->mode == KX_DEC_PLAIN) {
        if (!g->out || g->ld_out < N) { set_error("kx_decode_linear: KX_DEC_PLAIN needs out with ld_out >= N"); return KX_ERR_ARG; }
        p.out = g->out; p.ld_out = g->ld_out; p.out_f32 = g->out_f32; p.argmax_keys = g->argmax_keys;
    } else if (g->mode == KX_DEC_RESIDUAL) {
        if (!g->x || !g->xb || g->ld_x < N || g->ld_xb < N) { set_error("kx_decode_linear: KX_DEC_RESIDUAL needs x and xb"); return KX_ERR_ARG; }
        p.x = g->x; p.ld_x = g->ld_x; p.xb = reinterpret_cast<__nv_bfloat16*>(g->xb); p.ld_xb = g->ld_xb;
    } else if (g->mode == KX_DEC_QKV) {
        if (!g->q_out || !g->k_cache || !g->v_cache || !g->pos || !g->xq_cos || !g->xq_sin || !g->xk_cos || !g->xk_sin ||
            g->d_model <= 0 || N != 3 * g->d_model || (g->d_model & 63) || g->t_max <= 0 || g->ld_q < g->d_model) {
            set_error("kx_decode_linear: KX_DEC_QKV needs q_out, caches, pos, four xPos tables and N == 3*d_model (d_model %% 64 == 0)");
            return KX_ERR_ARG;
        }
        p.q_out = reinterpret_cast<__nv_bfloat16*>(g->q_out); p.ld_q = g->ld_q;
        p.k_cache = reinterpret_cast<__nv_bfloat16*>(g->k_cache); p.v_cache = reinterpret_cast<__nv_bfloat16*>(g->v_cache);
        p.t_max = g->t_max; p.d_model = g->d_model; p.pos = g->pos;
        p.xq_cos = g->xq_cos; p.xq_sin = g->xq_sin; p.xk_cos = g->xk_cos; p.xk_sin = g->xk_sin;
    } else {
        set_error("kx_decode_linear: unknown mode %d", g->mode);
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const int tiles = (N + 15) / 16;
    cudaError_t e;
    if (batch <= 8) {
        if (tiles <= sms + sms / 4) e = launch_chain(decode_linear_kernel<1, 8, 1>, dim3(tiles), dim3(256), stream, p);
        else e = launch_chain(decode_linear_kernel<1, 4, 3>, dim3(tiles), dim3(256), stream, p);
    } else if (batch <= 16) {
        e = launch_chain(decode_linear_kernel<2, 4, 2>, dim3(tiles), dim3(256), stream, p);
    } else {
        e = launch_chain(decode_linear_kernel<4, 4, 1>, dim3(tiles), dim3(256), stream, p);
    }
    return check_chain(e, "kx_decode_linear");
}

extern "C" size_t kx_decode_attn_scratch_bytes(int batch, int heads, int t_max) {
    if (batch <= 0 || heads <= 0 || t_max <= 0) return 0;
    const size_t chunks = (static_cast<size_t>(t_max) + DA_CHUNK - 1) / DA_CHUNK;
    return static_cast<size_t>(batch) * heads * chunks * 66 * sizeof(float);
}

extern "C" int kx_decode_attn(const void* q, long long ld_q, const void* k_cache, const void* v_cache, int t_max, int batch,
                              int heads, const int* pos, float scale, float* scratch, int* counters, void* out,
                              long long ld_out, cudaStream_t stream) {
    if (!q || !k_cache || !v_cache || !pos || !scratch || !counters || !out || batch <= 0 || heads <= 0 || t_max <= 0 ||
        (ld_q & 7) || ((uintptr_t)q & 15) || ((uintptr_t)k_cache & 15) || ((uintptr_t)v_cache & 15) || batch > 65535 ||
        heads > 65535) {
        set_error("kx_decode_attn: bad argument");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    const int chunks = (t_max + DA_CHUNK - 1) / DA_CHUNK;
    return check_chain(launch_chain(decode_attention_kernel, dim3(chunks, heads, batch), dim3(128), stream,
                                    reinterpret_cast<const __nv_bfloat16*>(q), ld_q, reinterpret_cast<const __nv_bfloat16*>(k_cache),
                                    reinterpret_cast<const __nv_bfloat16*>(v_cache), t_max, heads * 64, pos,
                                    scale * 1.4426950408889634f, scratch, counters, reinterpret_cast<__nv_bfloat16*>(out), ld_out,
                                    chunks), "kx_decode_attn");
}

extern "C" int kx_kv_cache_store(const void* qkv, long long ld_qkv, int batch, int seq_len, int d_model, void* k_cache,
                                 void* v_cache, int t_max, cudaStream_t stream) {
    if (!qkv || !k_cache || !v_cache || batch <= 0 || seq_len <= 0 || seq_len > t_max || d_model <= 0 || (d_model & 63) ||
        (ld_qkv & 7) || ld_qkv < 3LL * d_model || ((uintptr_t)qkv & 15) || ((uintptr_t)k_cache & 15) || ((uintptr_t)v_cache & 15)) {
        set_error("kx_kv_cache_store: bad argument (seq_len %d, t_max %d, d_model %d)", seq_len, t_max, d_model);
        return KX_ERR_ARG;
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const long long total = static_cast<long long>(batch) * seq_len * (d_model >> 3) * 2;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sms) * 8));
    kv_cache_store_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), ld_qkv, batch, seq_len,
                                                      d_model, reinterpret_cast<__nv_bfloat16*>(k_cache),
                                                      reinterpret_cast<__nv_bfloat16*>(v_cache), t_max);
    return check_launch("kx_kv_cache_store");
}

extern "C" int kx_decode_embed(const long long* tokens, int batch, const float* embed_table, int vocab, const float* pos_table,
                               int pos_rows, const int* pos, int text_index_off, int dim, float* x, void* xb, int* err_flag,
                               cudaStream_t stream) {
    if (!tokens || !embed_table || !pos_table || !pos || !x || !xb || batch <= 0 || vocab <= 0 || pos_rows <= 2 || dim <= 0 ||
        (dim & 3) || ((uintptr_t)embed_table & 15) || ((uintptr_t)pos_table & 15) || ((uintptr_t)x & 15) || ((uintptr_t)xb & 7)) {
        set_error("kx_decode_embed: bad argument");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    return check_chain(launch_chain(decode_embed_kernel, dim3(batch), dim3(256), stream, tokens, embed_table, vocab, pos_table,
                                    pos_rows, pos, text_index_off, dim, x, reinterpret_cast<__nv_bfloat16*>(xb), err_flag),
                       "kx_decode_embed");
}

extern "C" int kx_argmax_advance(const float* logits, long long ld, int batch, int vocab, const long long* forced,
                                 long long* tokens_out, long long* history, int history_ld, int* pos, int* step,
                                 int* counter, unsigned long long* argmax_keys, cudaStream_t stream) {
    if ((!logits && !argmax_keys) || !tokens_out || !step || !counter || batch <= 0 || vocab <= 0 || ld < vocab || (history && history_ld <= 0) ||
        (forced && history_ld <= 0)) {
        set_error("kx_argmax_advance: bad argument");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    return check_chain(launch_chain(argmax_advance_kernel, dim3(argmax_keys ? 1 : batch), dim3(argmax_keys ? 32 : 256), stream, logits,
                                    ld, vocab, forced, tokens_out, history, history_ld, pos, step, counter, argmax_keys, batch),
                       "kx_argmax_advance");
}
