// One decoding step as ONE persistent cooperative kernel.
//
// The per-kernel path (decode.cu) spends more time between kernels than inside them: a step is 123 launches of 5-35 MB
// each, and at 6.5 TB/s a 25 MB weight matrix is 4 us of work next to a launch + drain + first-load latency chain of
// about the same length.  Here the whole step — embedding, 24 x (q|k|v, attention, out_proj, fc1, fc2), LM head, greedy
// choice — runs in one grid of 3 CTAs per SM.  Phases are separated by a grid barrier (one 64-bit arrival counter,
// never reset: the target of barrier i in launch e is (e * barriers + i + 1) * gridDim.x), and before a CTA waits at a
// barrier it has already issued the first weight loads of its first work item of the NEXT phase — weights do not depend
// on the previous phase, so HBM stays busy while the barrier drains.  Work items: a Linear phase is cut into 16-row
// weight tiles (x split-K for the narrow out_proj / fc2 so that enough bytes are in flight; the last split to arrive
// sums the partials in a fixed order and runs the epilogue), the attention phase into (batch, head, 256-key chunk)
// items merged by the last arriver.  The tile arithmetic is the same as decode_linear_kernel / decode_attention_kernel
// (same MMA k-permutation, same folded LayerNorm, same fixed summation order within a tile).
//
// Activations written in one phase are read in the next by other SMs: they are read with ld.global.cg (L2), never
// through the non-coherent L1; weights and tables are immutable and use the streaming / read-only paths.
#include "kx_internal.h"
#include "ptx.cuh"

#include <vector>

namespace kx {

namespace {

enum { PH_EMBED = 0, PH_LINEAR = 1, PH_ATTN = 2, PH_PICK = 3 };
constexpr int STEP_THREADS = 256;
// Two register / occupancy trade-offs of the same kernel (KX_DECODE_STEP_VARIANT selects; default 0):
//   variant 0: 3 CTAs/SM, SU = 4 k-steps (8 x 16-byte weight loads) in flight per thread   (<= 80 registers)
//   variant 1: 2 CTAs/SM, SU = 8                                                             (<= 128 registers)
constexpr int ATTN_CHUNK = 256;             // keys per attention item (8 warps x 32)
constexpr int PART_STRIDE = 16 * 8 + 8 * 2; // split-K partial: 16x8 sums + 8 (sum, sumsq) pairs
constexpr int SPLITK_MAX_ITEMS = 2048;      // split phases have items <= grid (pick_ksplit), grid <= 3 * SMs

struct Phase {
    int type, items;
    // Linear
    const __nv_bfloat16* a; long long lda;
    const __nv_bfloat16* w; long long ldw; int N, K;
    const float* ln_c; const float* bias;
    int mode, act, ksplit;
    void* out; long long ld_out; int out_f32; unsigned long long* argmax_keys;
    // attention: q_out = q, k_cache, v_cache, out = attention output (bf16, ld_out)
    __nv_bfloat16* q_out; __nv_bfloat16* k_cache; __nv_bfloat16* v_cache;
};

struct StepCommon {
    int batch, d_model, heads, t_max, vocab, pos_rows, hist_ld, n_phases, attn_chunks;
    float eps, scale_log2;
    const long long* forced; long long* tokens; long long* history;
    const float* embed_table; const float* pos_table;
    const float *xq_cos, *xq_sin, *xk_cos, *xk_sin;
    float* x; __nv_bfloat16* xb;
    int* pos; int* step; int* err_flag;
    unsigned long long* argmax_keys;
    float* attn_part; int* attn_counters;
    float* splitk_part; int* splitk_counters;
    unsigned long long* barrier;            // [0] arrivals (monotonic), [1] launches completed, [16] release flag
    long long* trace;                       // optional: CTA 0 stamps globaltimer (work done, barrier left) per phase
};

struct StepPlan {
    StepCommon c;
    Phase ph[1];                            // n_phases entries
};

struct StepSmem {
    float red[8][16][8];
    float st[8][8][2];
    float fin[16][8];
    float att_o[8][64];
    float att_ml[8][2];
    int flag;
};

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void stats2(uint32_t v, float& s1, float& s2) {
    const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
    s1 += lo + hi;
    s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
}
__device__ __forceinline__ uint4 ldw_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_cg4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_cg_i(const int* p) {
    int v;
    asm volatile("ld.global.cg.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of the (cooperative, co-resident) grid arrive; a bounded spin so that a logic error cannot hang the GPU.
// Arrivals are counted on one line; the last arriver publishes the barrier's target on ANOTHER line (ctr[16]) that the
// waiters poll, so the polls do not queue behind the arrival atomics.
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target, int* err_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(ctr, 1ull);
        if (old + 1 == target) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;\n" :: "l"(ctr + 16), "l"(target) : "memory");
        } else {
            const long long t0 = clock64();
            while (ld_acquire_u64(ctr + 16) < target) {
                if (clock64() - t0 > (3ll << 30)) {         // ~2 s at 1.7 GHz
                    if (err_flag != nullptr) atomicOr(err_flag, 4);
                    break;
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

struct LinItem {
    int tile, split, s_begin, s_end;
    const uint4 *w0, *w1;
};

__device__ __forceinline__ LinItem lin_item(const Phase& p, int item) {
    LinItem it;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    it.tile = item / p.ksplit;
    it.split = item - it.tile * p.ksplit;
    const int sps = (p.K >> 5) / p.ksplit;
    const int spw = (sps + 7) >> 3;
    const int s0 = it.split * sps;
    it.s_begin = s0 + warp * spw;
    it.s_end = min(s0 + sps, it.s_begin + spw);
    const int n0 = it.tile * 16;
    const int r0 = min(n0 + g, p.N - 1), r1 = min(n0 + g + 8, p.N - 1);
    it.w0 = reinterpret_cast<const uint4*>(p.w + static_cast<long long>(r0) * p.ldw) + t;
    it.w1 = reinterpret_cast<const uint4*>(p.w + static_cast<long long>(r1) * p.ldw) + t;
    return it;
}

template <int SU>
__device__ __forceinline__ void lin_load_w(const LinItem& it, int s, uint4 (&wa)[SU], uint4 (&wb)[SU]) {
#pragma unroll
    for (int u = 0; u < SU; ++u) {
        const bool ok = s + u < it.s_end;
        wa[u] = ok ? ldw_stream(it.w0 + (s + u) * 4) : make_uint4(0, 0, 0, 0);
        wb[u] = ok ? ldw_stream(it.w1 + (s + u) * 4) : make_uint4(0, 0, 0, 0);
    }
}

// One (tile, split) item of a Linear phase.  `pre`: wa / wb already hold the first batch (loaded before the barrier).
template <int SU>
__device__ __forceinline__ void linear_item(const Phase& p, const StepCommon& C, int item, bool pre, uint4 (&wa)[SU],
                                            uint4 (&wb)[SU], StepSmem& sm) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const LinItem it = lin_item(p, item);
    const bool aok = g < C.batch;
    const uint4* ap = reinterpret_cast<const uint4*>(p.a + static_cast<long long>(aok ? g : 0) * p.lda) + t;
    const bool ln = p.ln_c != nullptr;

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float s1 = 0.f, s2 = 0.f;
    for (int s = it.s_begin; s < it.s_end; s += SU) {
        if (!pre) lin_load_w(it, s, wa, wb);
        pre = false;
        uint4 av[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u)
            av[u] = (aok && s + u < it.s_end) ? ld_cg4(ap + (s + u) * 4) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            mma16816(acc, wa[u].x, wb[u].x, wa[u].y, wb[u].y, av[u].x, av[u].y);
            mma16816(acc, wa[u].z, wb[u].z, wa[u].w, wb[u].w, av[u].z, av[u].w);
            if (ln) { stats2(av[u].x, s1, s2); stats2(av[u].y, s1, s2); stats2(av[u].z, s1, s2); stats2(av[u].w, s1, s2); }
        }
    }
    sm.red[warp][g][2 * t] = acc[0];
    sm.red[warp][g][2 * t + 1] = acc[1];
    sm.red[warp][g + 8][2 * t] = acc[2];
    sm.red[warp][g + 8][2 * t + 1] = acc[3];
    if (ln) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        if (t == 0) { sm.st[warp][g][0] = s1; sm.st[warp][g][1] = s2; }
    }
    __syncthreads();

    const int b = tid >> 4, r = tid & 15;                   // threads 0..127: (batch row b, feature r)
    float v = 0.f, a1 = 0.f, a2 = 0.f;
    if (tid < 128) {
#pragma unroll
        for (int w = 0; w < 8; ++w) v += sm.red[w][r][b];
        if (ln) {
#pragma unroll
            for (int w = 0; w < 8; ++w) { a1 += sm.st[w][b][0]; a2 += sm.st[w][b][1]; }
        }
    }
    if (p.ksplit > 1) {                                     // exchange split-K partials; the last arriver continues
        float* mine = C.splitk_part + static_cast<long long>(item) * PART_STRIDE;
        if (tid < 128) {
            mine[r * 8 + b] = v;
            if (r == 0) { mine[128 + 2 * b] = a1; mine[128 + 2 * b + 1] = a2; }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int done = atomicAdd(C.splitk_counters + it.tile, 1);
            sm.flag = (done == p.ksplit - 1);
            if (sm.flag) C.splitk_counters[it.tile] = 0;
        }
        __syncthreads();
        if (!sm.flag) return;
        __threadfence();
        if (tid < 128) {
            v = 0.f; a1 = 0.f; a2 = 0.f;
            const float* base = C.splitk_part + static_cast<long long>(it.tile) * p.ksplit * PART_STRIDE;
            for (int sp = 0; sp < p.ksplit; ++sp) {         // fixed order: bit-reproducible
                v += __ldcg(base + sp * PART_STRIDE + r * 8 + b);
                a1 += __ldcg(base + sp * PART_STRIDE + 128 + 2 * b);
                a2 += __ldcg(base + sp * PART_STRIDE + 128 + 2 * b + 1);
            }
        }
    }
    const int n0 = it.tile * 16;
    if (tid < 128) {
        const int n = min(n0 + r, p.N - 1);
        if (ln) {
            const float inv_n = 1.0f / static_cast<float>(p.K);
            const float mean = a1 * inv_n;
            const float var = fmaxf(a2 * inv_n - mean * mean, 0.f);
            const float rstd = rsqrtf(var + C.eps);
            v = fmaf(-mean * rstd, __ldg(p.ln_c + n), v * rstd);
        }
        if (p.bias != nullptr) v += __ldg(p.bias + n);
        sm.fin[r][b] = v;
    }
    __syncthreads();

    if (tid < 128) {
        const int n = n0 + r;
        const bool live = b < C.batch && n < p.N;
        v = sm.fin[r][b];
        if (p.argmax_keys != nullptr) {
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
            if (r == 0 && b < C.batch) atomicMax(p.argmax_keys + b, key);
        }
        if (live) {
            if (p.mode == KX_DEC_QKV) {
                const int which = n / C.d_model;
                const int col = n - which * C.d_model;
                const int pos = ld_cg_i(C.pos);
                if (which < 2) {
                    const int j = (n & 63) >> 1;
                    const float c = __ldg((which == 0 ? C.xq_cos : C.xk_cos) + pos * 32 + j);
                    const float s = __ldg((which == 0 ? C.xq_sin : C.xk_sin) + pos * 32 + j);
                    const float x0 = sm.fin[r & ~1][b], x1 = sm.fin[r | 1][b];
                    v = (r & 1) ? fmaf(x1, c, x0 * s) : fmaf(x0, c, -(x1 * s));
                }
                if (which == 0) {
                    p.q_out[static_cast<long long>(b) * C.d_model + col] = __float2bfloat16_rn(v);
                } else if (pos < C.t_max) {
                    __nv_bfloat16* dst = (which == 1 ? p.k_cache : p.v_cache);
                    dst[((static_cast<long long>(b) * C.heads + (col >> 6)) * C.t_max + pos) * 64 + (col & 63)] =
                        __float2bfloat16_rn(v);
                }
            } else if (p.mode == KX_DEC_RESIDUAL) {
                float* px = C.x + static_cast<long long>(b) * C.d_model + n;
                v += __ldcg(px);
                *px = v;
                C.xb[static_cast<long long>(b) * C.d_model + n] = __float2bfloat16_rn(v);
            } else {
                if (p.act == KX_ACT_GELU) v = gelu_erf(v);
                else if (p.act == KX_ACT_QUICK_GELU) v = quick_gelu(v);
                if (p.out_f32) reinterpret_cast<float*>(p.out)[static_cast<long long>(b) * p.ld_out + n] = v;
                else reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(b) * p.ld_out + n] = __float2bfloat16_rn(v);
            }
        }
    }
    // the next item of this CTA writes sm.red only after its own k loop and a __syncthreads: no extra barrier needed,
    // but sm.fin is read above and rewritten after that barrier, and sm.flag after two: safe.
}

// One (batch, head, 256-key chunk) item of the attention phase.
__device__ __forceinline__ void attn_item(const Phase& p, const StepCommon& C, int item, StepSmem& sm) {
    const int H = C.heads, chunks = C.attn_chunks;
    const int chunk = item % chunks;
    const int bh = item / chunks;
    const int b = bh / H, h = bh - b * H;
    const int n_keys = min(ld_cg_i(C.pos) + 1, C.t_max);
    const int n_act = (n_keys + ATTN_CHUNK - 1) / ATTN_CHUNK;
    if (chunk >= n_act) return;                              // uniform over the CTA
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & 7, kq = lane >> 3;

    float qf[8];
    {
        const uint4 raw = ld_cg4(reinterpret_cast<const uint4*>(p.q_out + static_cast<long long>(b) * C.d_model + h * 64) + sub);
        const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            qf[2 * i] = __uint_as_float(r[i] << 16) * C.scale_log2;
            qf[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u) * C.scale_log2;
        }
    }
    const int key0 = chunk * ATTN_CHUNK + warp * 32 + kq;
    const long long base = (static_cast<long long>(b) * H + h) * C.t_max * 64;
    float sc[8];
    float m = -INFINITY;
    {
        uint4 kr[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = key0 + it * 4;
            kr[it] = j < n_keys ? ld_cg4(reinterpret_cast<const uint4*>(p.k_cache + base + static_cast<long long>(j) * 64) + sub)
                                : make_uint4(0, 0, 0, 0);
        }
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
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    float l = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (m > -INFINITY) {
        uint4 vr[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = key0 + it * 4;
            vr[it] = j < n_keys ? ld_cg4(reinterpret_cast<const uint4*>(p.v_cache + base + static_cast<long long>(j) * 64) + sub)
                                : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const float pj = ex2_approx(sc[it] - m);
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
    __syncthreads();                                         // previous item's readers of att_o / att_ml are done
    if (kq == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.att_o[warp][sub * 8 + i] = o[i];
        if (sub == 0) { sm.att_ml[warp][0] = m; sm.att_ml[warp][1] = l; }
    }
    __syncthreads();
    float* mine = C.attn_part + (static_cast<long long>(bh) * chunks + chunk) * 66;
    if (threadIdx.x < 64) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, sm.att_ml[w][0]);
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float f = sm.att_ml[w][0] > -INFINITY ? ex2_approx(sm.att_ml[w][0] - M) : 0.f;
            L = fmaf(sm.att_ml[w][1], f, L);
            O = fmaf(sm.att_o[w][threadIdx.x], f, O);
        }
        if (n_act == 1) {                                    // single chunk: no exchange
            reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(b) * p.ld_out + h * 64 + threadIdx.x] =
                __float2bfloat16_rn(O / L);
        } else {
            mine[threadIdx.x] = O;
            if (threadIdx.x == 0) { mine[64] = M; mine[65] = L; }
        }
    }
    if (n_act == 1) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(C.attn_counters + bh, 1);
        sm.flag = (done == n_act - 1);
        if (sm.flag) C.attn_counters[bh] = 0;
    }
    __syncthreads();
    if (!sm.flag) return;
    __threadfence();
    if (threadIdx.x < 64) {
        const float* all = C.attn_part + static_cast<long long>(bh) * chunks * 66;
        float M = -INFINITY, L = 0.f, O = 0.f;
        for (int c0 = 0; c0 < n_act; c0 += 8) {
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
        reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(b) * p.ld_out + h * 64 + threadIdx.x] =
            __float2bfloat16_rn(O / L);
    }
}

__device__ __forceinline__ void embed_row(const StepCommon& C, int b) {
    long long id = C.tokens[b];
    if (id < 0 || id >= C.vocab) {
        if (threadIdx.x == 0 && C.err_flag != nullptr) atomicOr(C.err_flag, 1);
        id = 0;
    }
    int pr = ld_cg_i(C.pos) + 2;
    if (pr >= C.pos_rows) {
        if (threadIdx.x == 0 && C.err_flag != nullptr) atomicOr(C.err_flag, 2);
        pr = C.pos_rows - 1;
    }
    const float4* e = reinterpret_cast<const float4*>(C.embed_table + id * C.d_model);
    const float4* pp = reinterpret_cast<const float4*>(C.pos_table + static_cast<long long>(pr) * C.d_model);
    for (int i = threadIdx.x; i < (C.d_model >> 2); i += blockDim.x) {
        const float4 a = __ldg(e + i), c = __ldg(pp + i);
        const float4 r = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
        reinterpret_cast<float4*>(C.x + static_cast<long long>(b) * C.d_model)[i] = r;
        uint2 pk;
        pk.x = pack_bf16(r.x, r.y); pk.y = pack_bf16(r.z, r.w);
        reinterpret_cast<uint2*>(C.xb + static_cast<long long>(b) * C.d_model)[i] = pk;
    }
}

__device__ __forceinline__ void pick_tokens(const StepCommon& C) {
    const int step = ld_cg_i(C.step);
    __syncthreads();
    for (int b = threadIdx.x; b < C.batch; b += blockDim.x) {
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(C.argmax_keys + b);
        long long choice = static_cast<long long>(0xffffffffu - static_cast<uint32_t>(key & 0xffffffffull));
        C.argmax_keys[b] = 0ull;
        if (C.forced != nullptr && step < C.hist_ld) choice = C.forced[static_cast<long long>(b) * C.hist_ld + step];
        C.tokens[b] = choice;
        if (C.history != nullptr && step < C.hist_ld) C.history[static_cast<long long>(b) * C.hist_ld + step] = choice;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *C.step = step + 1;
        *C.pos = ld_cg_i(C.pos) + 1;
    }
}

}  // namespace

template <int CPS, int SU>
__global__ void __launch_bounds__(STEP_THREADS, CPS)
decode_step_kernel(const StepPlan* __restrict__ plan) {
    __shared__ StepSmem sm;
    const StepCommon& C = plan->c;
    const int n_phases = C.n_phases;
    unsigned long long* bar = C.barrier;
    const unsigned long long epoch = ld_acquire_u64(bar + 1);
    const unsigned long long per_launch = static_cast<unsigned long long>(n_phases - 1) * gridDim.x;
    unsigned long long target = epoch * per_launch;

    uint4 wa[SU], wb[SU];
    bool pre = false;
    for (int ph = 0; ph < n_phases; ++ph) {
        const Phase& P = plan->ph[ph];
        const int type = P.type, items = P.items;
        if (type == PH_LINEAR) {
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                linear_item(P, C, item, pre, wa, wb, sm);
                pre = false;
            }
        } else if (type == PH_ATTN) {
            for (int item = blockIdx.x; item < items; item += gridDim.x) attn_item(P, C, item, sm);
        } else if (type == PH_EMBED) {
            if (static_cast<int>(blockIdx.x) < C.batch) embed_row(C, blockIdx.x);
        } else {
            if (blockIdx.x == 0) pick_tokens(C);
        }
        pre = false;
        if (C.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) C.trace[2 * ph] = global_ns();
        if (ph + 1 < n_phases) {
            const Phase& Nx = plan->ph[ph + 1];
            if (Nx.type == PH_LINEAR && static_cast<int>(blockIdx.x) < Nx.items) {
                const LinItem it = lin_item(Nx, blockIdx.x);  // weights are immutable: stream them in while the barrier drains
                lin_load_w(it, it.s_begin, wa, wb);
                pre = true;
            }
            target += gridDim.x;
            grid_barrier(bar, target, C.err_flag);
            if (C.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) C.trace[2 * ph + 1] = global_ns();
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence();
        bar[1] = epoch + 1;
    }
}

namespace {

int step_variant() {
    static const int v = [] {
        const char* e = getenv("KX_DECODE_STEP_VARIANT");
        return (e != nullptr && e[0] == '1') ? 1 : 0;
    }();
    return v;
}

int step_grid(int sms) {
    static int per_sm = -1;
    if (per_sm < 0) {
        int n = 0;
        const cudaError_t e = step_variant() == 1
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_step_kernel<2, 8>, STEP_THREADS, 0)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_step_kernel<3, 4>, STEP_THREADS, 0);
        if (e != cudaSuccess) n = 0;
        per_sm = std::min(n, step_variant() == 1 ? 2 : 3);
    }
    return per_sm * sms;
}

int pick_ksplit(int tiles, int steps, int grid) {
    int ks = 1;
    while (ks < 8 && tiles * ks * 2 <= grid && steps % (ks * 2) == 0 && steps / (ks * 2) >= 8) ks *= 2;
    return ks;
}

}  // namespace

}  // namespace kx

using namespace kx;

static int step_phase_count(int layers) { return 1 + 5 * layers + 2; }

extern "C" size_t kx_decode_plan_bytes(int layers) {
    if (layers < 0) return 0;
    return sizeof(StepPlan) + sizeof(Phase) * static_cast<size_t>(step_phase_count(layers));
}

extern "C" size_t kx_decode_step_scratch_floats(int batch, int heads, int t_max) {
    if (batch <= 0 || heads <= 0 || t_max <= 0) return 0;
    const size_t chunks = (static_cast<size_t>(t_max) + ATTN_CHUNK - 1) / ATTN_CHUNK;
    return static_cast<size_t>(batch) * heads * chunks * 66 + static_cast<size_t>(SPLITK_MAX_ITEMS) * PART_STRIDE;
}

extern "C" size_t kx_decode_step_counters(int batch, int heads) {
    if (batch <= 0 || heads <= 0) return 0;
    return static_cast<size_t>(batch) * heads + SPLITK_MAX_ITEMS;
}

extern "C" int kx_decode_plan_build(const kx_decode_step_args* g, void* device_plan, cudaStream_t stream) {
    if (!g || !device_plan) { set_error("kx_decode_plan_build: null argument"); return KX_ERR_ARG; }
    if (g->batch <= 0 || g->batch > 8 || g->layers < 0 || g->d_model <= 0 || (g->d_model & 63) || g->ffn <= 0 || (g->ffn & 31) ||
        g->heads * 64 != g->d_model || g->vocab <= 0 || g->t_max <= 0 || g->pos_rows <= 2) {
        set_error("kx_decode_plan_build: need 1 <= batch <= 8, d_model == heads*64, ffn %% 32 == 0 (batch %d, d_model %d, heads %d)",
                  g->batch, g->d_model, g->heads);
        return KX_ERR_ARG;
    }
    const void* need[] = {g->w_qkv, g->c_qkv, g->d_qkv, g->w_o, g->c_o, g->d_o, g->w_fc1, g->c_fc1, g->d_fc1, g->w_fc2, g->c_fc2,
                          g->d_fc2, g->k_cache, g->v_cache, g->w_out, g->c_out, g->embed_table, g->pos_table, g->xq_cos, g->xq_sin,
                          g->xk_cos, g->xk_sin, g->tokens, g->x, g->xb, g->q, g->att, g->mid, g->logits, g->argmax_keys, g->pos,
                          g->step, g->scratch, g->counters, g->barrier};
    for (const void* q : need)
        if (q == nullptr) { set_error("kx_decode_plan_build: a required pointer is NULL"); return KX_ERR_ARG; }
    if ((g->forced || g->history) && g->history_ld <= 0) { set_error("kx_decode_plan_build: history_ld"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const int grid = step_grid(sms);
    if (grid <= 0) { set_error("kx_decode_plan_build: decode_step_kernel does not fit on an SM"); return KX_ERR_LAUNCH; }

    const int L = g->layers, D = g->d_model, F = g->ffn, n_ph = step_phase_count(L);
    const size_t bytes = kx_decode_plan_bytes(L);
    std::vector<unsigned char> host(bytes, 0);
    StepPlan* plan = reinterpret_cast<StepPlan*>(host.data());
    StepCommon& C = plan->c;
    const int chunks = (g->t_max + ATTN_CHUNK - 1) / ATTN_CHUNK;
    C.batch = g->batch; C.d_model = D; C.heads = g->heads; C.t_max = g->t_max; C.vocab = g->vocab; C.pos_rows = g->pos_rows;
    C.hist_ld = g->history_ld; C.n_phases = n_ph; C.attn_chunks = chunks;
    C.eps = g->eps; C.scale_log2 = g->scale * 1.4426950408889634f;
    C.forced = g->forced; C.tokens = g->tokens; C.history = g->history;
    C.embed_table = g->embed_table; C.pos_table = g->pos_table;
    C.xq_cos = g->xq_cos; C.xq_sin = g->xq_sin; C.xk_cos = g->xk_cos; C.xk_sin = g->xk_sin;
    C.x = g->x; C.xb = reinterpret_cast<__nv_bfloat16*>(g->xb);
    C.pos = g->pos; C.step = g->step; C.err_flag = g->err_flag; C.argmax_keys = g->argmax_keys;
    C.attn_part = g->scratch;
    C.splitk_part = g->scratch + static_cast<size_t>(g->batch) * g->heads * chunks * 66;
    C.attn_counters = g->counters;
    C.splitk_counters = g->counters + g->batch * g->heads;
    C.barrier = g->barrier;
    C.trace = g->trace;

    int k = 0;
    auto lin = [&](const void* a, long long lda, const void* w, int N, int K, const float* c, const float* d, int mode, int act,
                   bool allow_split) -> Phase& {
        Phase& P = plan->ph[k++];
        P.type = PH_LINEAR;
        P.a = reinterpret_cast<const __nv_bfloat16*>(a); P.lda = lda;
        P.w = reinterpret_cast<const __nv_bfloat16*>(w); P.ldw = K; P.N = N; P.K = K;
        P.ln_c = c; P.bias = d; P.mode = mode; P.act = act;
        const int tiles = (N + 15) / 16;
        P.ksplit = allow_split ? pick_ksplit(tiles, K >> 5, grid) : 1;
        P.items = tiles * P.ksplit;
        return P;
    };
    plan->ph[k].type = PH_EMBED; plan->ph[k].items = g->batch; ++k;
    for (int l = 0; l < L; ++l) {
        Phase& q = lin(g->xb, D, g->w_qkv[l], 3 * D, D, g->c_qkv[l], g->d_qkv[l], KX_DEC_QKV, KX_ACT_NONE, false);
        q.q_out = reinterpret_cast<__nv_bfloat16*>(g->q);
        q.k_cache = reinterpret_cast<__nv_bfloat16*>(g->k_cache[l]); q.v_cache = reinterpret_cast<__nv_bfloat16*>(g->v_cache[l]);
        Phase& at = plan->ph[k++];
        at.type = PH_ATTN; at.items = g->batch * g->heads * chunks;
        at.q_out = reinterpret_cast<__nv_bfloat16*>(g->q);
        at.k_cache = reinterpret_cast<__nv_bfloat16*>(g->k_cache[l]); at.v_cache = reinterpret_cast<__nv_bfloat16*>(g->v_cache[l]);
        at.out = g->att; at.ld_out = D;
        lin(g->att, D, g->w_o[l], D, D, g->c_o[l], g->d_o[l], KX_DEC_RESIDUAL, KX_ACT_NONE, true);
        Phase& f1 = lin(g->xb, D, g->w_fc1[l], F, D, g->c_fc1[l], g->d_fc1[l], KX_DEC_PLAIN, KX_ACT_GELU, false);
        f1.out = g->mid; f1.ld_out = F; f1.out_f32 = 0;
        lin(g->mid, F, g->w_fc2[l], D, F, g->c_fc2[l], g->d_fc2[l], KX_DEC_RESIDUAL, KX_ACT_NONE, true);
    }
    Phase& hd = lin(g->xb, D, g->w_out, g->vocab, D, g->c_out, g->d_out, KX_DEC_PLAIN, KX_ACT_NONE, false);
    hd.out = g->logits; hd.ld_out = g->ld_logits; hd.out_f32 = 1; hd.argmax_keys = g->argmax_keys;
    plan->ph[k].type = PH_PICK; plan->ph[k].items = 1; ++k;
    for (int i = 0; i < k; ++i)
        if (plan->ph[i].type == PH_LINEAR && ((plan->ph[i].K & 31) || (plan->ph[i].ksplit > 1 && plan->ph[i].items > SPLITK_MAX_ITEMS))) {
            set_error("kx_decode_plan_build: unsupported Linear shape (K %d)", plan->ph[i].K);
            return KX_ERR_ARG;
        }
    // one-time setup call: the plan is copied from a temporary host buffer, so the stream is synchronised here
    cudaError_t e = cudaMemcpyAsync(device_plan, host.data(), bytes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("kx_decode_plan_build: %s", cudaGetErrorString(e)); return KX_ERR_LAUNCH; }
    return KX_OK;
}

extern "C" int kx_decode_step(const void* device_plan, cudaStream_t stream) {
    if (!device_plan) { set_error("kx_decode_step: null plan"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const int grid = step_grid(sms);
    if (grid <= 0) { set_error("kx_decode_step: decode_step_kernel does not fit on an SM"); return KX_ERR_LAUNCH; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(STEP_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;              // co-residency of the whole grid is what the barrier relies on
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const StepPlan* plan = reinterpret_cast<const StepPlan*>(device_plan);
    const cudaError_t e = step_variant() == 1 ? cudaLaunchKernelEx(&cfg, decode_step_kernel<2, 8>, plan)
                                              : cudaLaunchKernelEx(&cfg, decode_step_kernel<3, 4>, plan);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("kx_decode_step launch failed: %s", cudaGetErrorString(e));
        return KX_ERR_LAUNCH;
    }
    count_launch();
    return KX_OK;
}
