// One decoding step as ONE persistent cooperative kernel.
//
// A decoding step streams 2.55 GB of weights plus the KV cache and does almost no arithmetic: the only thing that
// matters is that HBM never idles.  The per-kernel path (decode.cu) cannot keep it busy: 123 launches of 5-35 MB each
// are latency-bound streams (ncu: long-scoreboard stalls, 24 warps/SM) separated by launch / drain gaps.  Here the whole
// step — embedding, 24 x (q|k|v, attention, out_proj, fc1, fc2), LM head, greedy choice — is one grid of ONE CTA (9 warps)
// per SM (two do not fit: 3 + 3 warps of 112 registers exceed a 16 K-register sub-partition — found in the ncu launch table):
//
//   * warp 8 of every CTA is a PRODUCER: it walks the CTA's whole work list of the step and moves it through a
//     6-stage, 33 KB-per-stage (200 KB) shared-memory ring with bulk copies (cp.async.bulk, mbarrier complete_tx; lane 0 arms the
//     barrier, 16 lanes issue the row copies at once): 16 weight rows x 1024 k per stage for a Linear item, 256 keys
//     of K or of V for an attention item.  Weights and the history
//     rows of the cache are immutable during the launch, so the producer never waits for a phase boundary: while the
//     consumers of a CTA sit in a grid barrier its ring is already filling with the NEXT phase's bytes (200 KB per SM
//     in flight, several times what Little's law asks for at 6.5 TB/s; a whole phase's share of a CTA fits the ring).
//   * warps 0-7 are CONSUMERS: mma.sync m16n8k16 with the weight rows as the M dimension and the (<= 8) sequences as N,
//     fragments read from the ring with conflict-free 16-byte LDS (rows padded by 64 bytes), the same k-permutation,
//     folded LayerNorm, fixed-order reductions and epilogues as decode_linear_kernel; attention reads its 256-key chunk
//     from the ring and the single NEW key/value row (written by the q|k|v phase of this launch) straight from L2.
//   * phases are separated by a grid barrier over the consumers (arrivals on one line, release flag on another).
//     Every dependent L2 round trip inside a phase is on the critical path of the whole GPU, so the work is cut to
//     keep those chains short: a Linear phase is ONE pass (a CTA takes up to four 16-row tiles that share the
//     activation fragments; no split-K exchange), the epilogue operands are requested at item entry, an attention item
//     is a whole (batch, head) — its chunks are folded with an online softmax inside the CTA, no cross-CTA merge — and
//     the attention phase does not wait at a grid barrier at all: it waits for the 12 q|k|v tiles of ITS head.
//
// Activations written in one phase are read in the next by other SMs with ld.global.cg (L2), never through L1.
//
// Status (profiles/r1_decode_bench.md): bit-identical to the per-kernel path on the test model, 1.00 ms per step at B = 8
// against 1.03 ms for the CUDA graph of separate kernels (1.37 vs 1.41 ms at a 1920-row prompt): generate()'s default for
// batches of at most 8 sequences.
#include "kx_internal.h"
#include "ptx.cuh"

#include <vector>

namespace kx {

namespace {

enum { PH_EMBED = 0, PH_LINEAR = 1, PH_ATTN = 2, PH_PICK = 3 };
constexpr int CONSUMER_WARPS = 8;
constexpr int CONSUMERS = CONSUMER_WARPS * 32;
constexpr int STEP_THREADS = CONSUMERS + 32;
constexpr int STEP_CTAS_PER_SM = 1;
constexpr int NS = 6;                           // ring stages
constexpr int CHUNK_K = 1024;                   // k per weight stage
constexpr int W_PITCH = CHUNK_K * 2 + 64;       // bytes per staged weight row (+64: the quad-row LDS.128 pattern hits all banks)
constexpr int STAGE_BYTES = 16 * W_PITCH;       // 33792 >= 256 keys x 128 bytes
constexpr int ATTN_CHUNK = 256;                 // keys per attention item (8 warps x 32)
constexpr int HEAD_CNT0 = 32;                   // first per-head counter in the barrier buffer (uint64 index)
constexpr int TILES_PER_HEAD = 12;              // 64 features / 16 per tile, for each of q, k, v
static_assert(STAGE_BYTES >= ATTN_CHUNK * 128 && STAGE_BYTES % 128 == 0, "ring stage geometry");

struct Phase {
    int type, items;
    // Linear
    const __nv_bfloat16* a; long long lda;
    const __nv_bfloat16* w; long long ldw; int N, K;
    const float* ln_c; const float* bias;
    int mode, act, nt;                       // nt: weight tiles per item (1..4, sharing the activation fragments)
    int layer, no_barrier;                   // q|k|v phases: the attention phase behind them waits per head, not at a grid barrier
    void* out; long long ld_out; int out_f32; unsigned long long* argmax_keys;
    // attention: q_out = q, k_cache, v_cache, out = attention output (bf16, ld_out)
    __nv_bfloat16* q_out; __nv_bfloat16* k_cache; __nv_bfloat16* v_cache;
};

struct StepCommon {
    int batch, d_model, heads, t_max, vocab, pos_rows, hist_ld, n_phases;
    int text_off;                           // >= 0: second positional add at (*pos + 2 - text_off), see kx_decode_embed
    float eps, scale_log2;
    const long long* forced; long long* tokens; long long* history;
    const float* embed_table; const float* pos_table;
    const float *xq_cos, *xq_sin, *xk_cos, *xk_sin;
    float* x; __nv_bfloat16* xb;
    int* pos; int* step; int* err_flag;
    unsigned long long* argmax_keys;
    unsigned long long* barrier;            // [0] arrivals (monotonic), [1] launches completed, [16] release flag,
                                            // [HEAD_CNT0 + h] q|k|v tiles of head h stored so far (monotonic, 12 per layer)
    int layers;
    long long* trace;                       // optional: CTA 0 stamps globaltimer (work done, barrier left) per phase
};

struct StepPlan {
    StepCommon c;
    Phase ph[1];                            // n_phases entries
};

struct StepSmem {
    float red[4][8][16][8];
    float st[8][8][2];
    float fin[4][16][8];
    float att_o[8][64];
    float att_ml[8][2];
    uint64_t full[NS], empty[NS];
    StepCommon c;                           // the plan, staged: field reads are LDS, not dependent global loads
    Phase ph[2];                            // current / next phase (the next one is fetched before the barrier)
};

__device__ __forceinline__ void stage_words(void* dst_smem, const void* src_global, int bytes, int tid, int nthreads) {
    uint32_t* d = reinterpret_cast<uint32_t*>(dst_smem);
    const uint32_t* g = reinterpret_cast<const uint32_t*>(src_global);
    for (int i = tid; i < bytes / 4; i += nthreads) d[i] = __ldg(g + i);
}

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
__device__ __forceinline__ uint4 ld_cg4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
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
// Polling load: relaxed, so that ptxas does not attach an L1 invalidation (CCTL.IVALL) to every iteration — ncu showed
// the acquire version flushing L1 on each poll, which turned every later read of the plan into an L2 round trip.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// global -> shared bulk copy completing on an mbarrier (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { named_bar_sync(1, CONSUMERS); }
// mbarrier wait without a diagnostic printf (its argument block is a stack frame, and local memory is an L2 round trip
// in this kernel); a pipeline bug still traps instead of hanging.
__device__ __forceinline__ void ring_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (3ll << 31)) __trap();
    }
}
// fine-grained profiling stamps of CTA 0 / thread 0 inside one chosen phase (trace[2*n_phases + 1 + k]; the chosen phase
// index is read from trace[2*n_phases], -1 = none)
__device__ __forceinline__ void stamp(long long* dbg, int k) {
#ifdef KX_STEP_FINE_TRACE                   // costs registers: the kernel spills with it (and spills are L2 round trips here,
    if (dbg != nullptr) dbg[k] = static_cast<long long>(globaltimer_ns());     // L1 is invalidated at every grid barrier)
#else
    (void)dbg; (void)k;
#endif
}

// Measured alternatives: a separate release line written by the last arriver and polled by everybody (one more hop:
// +0.8 us per barrier, 1.08 vs 1.00 ms per step); a two-level variant with 8 shard counters (the second dependent atomic
// costs more than the contention it removes: 3.1-4.0 us of barrier wait per phase).
// Grid barrier over the consumers of all (co-resident) CTAs: one monotonic arrival counter, polled with relaxed loads by
// thread 0 of every CTA (one acquire at the end).  Bounded spin: a logic error cannot hang the GPU.
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target, int* err_flag, long long* dbg) {
    consumer_sync();
    if (threadIdx.x == 0) {
        stamp(dbg, 8);
        // release on the arrival itself (cumulative over the CTA's writes ordered before it by the bar.sync above) instead
        // of a stand-alone MEMBAR.SC: the full fence also waited for the producer's in-flight bulk copies to drain.
        // acq_rel, not release: the LAST arriver skips the polling loop below, and its arrival must itself synchronise-with
        // the other CTAs' releases (it is off the critical path: everybody else is still spinning on its increment).
        unsigned long long old;
        asm volatile("atom.add.acq_rel.gpu.global.u64 %0, [%1], 1;\n" : "=l"(old) : "l"(ctr) : "memory");
        stamp(dbg, 9);
        if (old + 1 != target) {                            // poll the arrival counter itself: one hop from the last arrival
            const long long t0 = clock64();
            while (ld_relaxed_u64(ctr) < target) {
                if (clock64() - t0 > (3ll << 30)) {         // ~2 s
                    if (err_flag != nullptr) atomicOr(err_flag, 4);
                    break;
                }
            }
            (void)ld_acquire_u64(ctr);                      // one acquire (one L1 invalidation) once the count is seen
        }
    }
    consumer_sync();
}

// ---- the work list of one CTA, walked identically by its producer and its consumers -------------------------------
struct Ring {
    uint32_t base;                          // shared address of stage 0
    uint64_t* full;
    uint64_t* empty;
    uint32_t i;                             // stages used so far
    __device__ __forceinline__ uint32_t slot() const { return i % NS; }
    __device__ __forceinline__ uint32_t parity() const { return (i / NS) & 1; }
    __device__ __forceinline__ uint32_t addr() const { return base + slot() * STAGE_BYTES; }
};

// Producer warp, all 32 lanes: lane 0 waits for the slot and arms the barrier, then lane `row` issues that row's copy —
// sixteen bulk copies leave in one warp instruction instead of sixteen dependent issues by one thread (measured: the
// single-thread version needed ~1.3 us per 32 KB stage, i.e. 25 GB/s per CTA).
__device__ __forceinline__ void produce_linear(const Phase& p, int item, Ring& r, int lane) {
    const int tiles = (p.N + 15) >> 4;
    const int K = p.K, N = p.N, nt = p.nt, items = p.items;
    const long long ldw = p.ldw;
    const __nv_bfloat16* w = p.w;
    for (int kc = 0; kc < K; kc += CHUNK_K) {
        const int ck = min(CHUNK_K, K - kc);
        for (int tt = 0; tt < nt; ++tt) {
            const int tile = item + tt * items;
            if (tile >= tiles) break;
            const int n0 = tile * 16;
            const int rows = min(16, N - n0);
            if (lane == 0) {
                ring_wait(r.empty + r.slot(), r.parity() ^ 1);
                mbar_arrive_expect_tx(r.full + r.slot(), static_cast<uint32_t>(rows * ck * 2));
            }
            __syncwarp();
            if (lane < rows)
                bulk_g2s(r.addr() + lane * W_PITCH, w + static_cast<long long>(n0 + lane) * ldw + kc, static_cast<uint32_t>(ck * 2),
                         r.full + r.slot());
            ++r.i;
        }
    }
}

// Attention item: the K and V chunks are contiguous 32 KB streams; 8 lanes x 4 KB each.
__device__ __forceinline__ void produce_attn(const Phase& p, const StepCommon& C, int bh, int n_keys, Ring& r, int lane) {
    const int n_chunks = (n_keys + ATTN_CHUNK - 1) / ATTN_CHUNK;
    const __nv_bfloat16* kc_ptr = p.k_cache;
    const __nv_bfloat16* vc_ptr = p.v_cache;
    for (int c = 0; c < n_chunks; ++c) {
        // history rows only (keys < n_keys - 1): the newest row is written by this launch and read from L2 by the consumers
        const int rows = min(n_keys - 1, (c + 1) * ATTN_CHUNK) - c * ATTN_CHUNK;
        const long long off = (static_cast<long long>(bh) * C.t_max + static_cast<long long>(c) * ATTN_CHUNK) * 64;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            if (lane == 0) {
                ring_wait(r.empty + r.slot(), r.parity() ^ 1);
                if (rows > 0) mbar_arrive_expect_tx(r.full + r.slot(), static_cast<uint32_t>(rows * 128));
                else mbar_arrive(r.full + r.slot());
            }
            __syncwarp();
            const int r_lo = lane * 32, r_n = min(32, rows - r_lo);          // 32 rows = 4 KB per lane
            if (lane < 8 && r_n > 0)
                bulk_g2s(r.addr() + r_lo * 128, (which ? vc_ptr : kc_ptr) + off + r_lo * 64, static_cast<uint32_t>(r_n * 128),
                         r.full + r.slot());
            ++r.i;
        }
    }
}

__device__ void producer(const StepPlan* plan, const StepCommon& C, Ring r) {
    const int lane = threadIdx.x & 31;
    const int n_keys = min(ld_cg_i(C.pos) + 1, C.t_max);     // pos only changes in the last phase of a launch
    const int n_phases = C.n_phases;
    for (int ph = 0; ph < n_phases; ++ph) {
        const Phase& P = plan->ph[ph];
        const int type = P.type, items = P.items;
        if (type == PH_LINEAR) {
            for (int item = blockIdx.x; item < items; item += gridDim.x) produce_linear(P, item, r, lane);
        } else if (type == PH_ATTN) {
            for (int item = blockIdx.x; item < items; item += gridDim.x) produce_attn(P, C, item, n_keys, r, lane);
        }
    }
}

// One item of a Linear phase (consumer warps): NT weight tiles against the same activation fragments.
template <int NT>
__device__ __forceinline__ void linear_item(const Phase& p, const StepCommon& C, int item, int pos0, Ring& r, StepSmem& sm,
                                            long long* dbg) {
    stamp(dbg, 0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int tiles = (p.N + 15) >> 4;
    const bool aok = g < C.batch;
    const uint4* ap = reinterpret_cast<const uint4*>(p.a + static_cast<long long>(aok ? g : 0) * p.lda) + t;
    const bool ln = p.ln_c != nullptr;
    bool tile_ok[NT];
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) tile_ok[tt] = item + tt * p.items < tiles;

    // Everything the epilogue of this thread's (tile, batch row, feature) needs is requested NOW, so that those DRAM /
    // L2 round trips overlap the k loop instead of sitting on the phase's critical path.
    constexpr int ROUNDS = (NT + 1) / 2;                    // 256 threads finish two tiles per round
    const int e_half = tid >> 7, e_idx = tid & 127;
    const int e_b = e_idx >> 4, e_r = e_idx & 15;
    float e_lnc[ROUNDS], e_bias[ROUNDS], e_c[ROUNDS], e_s[ROUNDS], e_x[ROUNDS];
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
        const int e_tt = 2 * rd + e_half;
        const bool e_mine = e_tt < NT && item + e_tt * p.items < tiles;
        const int e_n = min((item + e_tt * p.items) * 16 + e_r, p.N - 1);
        e_lnc[rd] = 0.f; e_bias[rd] = 0.f; e_c[rd] = 1.f; e_s[rd] = 0.f; e_x[rd] = 0.f;
        if (e_mine) {
            if (ln) e_lnc[rd] = __ldg(p.ln_c + e_n);
            if (p.bias != nullptr) e_bias[rd] = __ldg(p.bias + e_n);
            if (p.mode == KX_DEC_QKV) {
                const int which = e_n / C.d_model;
                if (which < 2) {                            // pos0 was read once at kernel start: no dependent round trip here
                    const int j = (e_n & 63) >> 1;
                    e_c[rd] = __ldg((which == 0 ? C.xq_cos : C.xk_cos) + pos0 * 32 + j);
                    e_s[rd] = __ldg((which == 0 ? C.xq_sin : C.xk_sin) + pos0 * 32 + j);
                }
            } else if (p.mode == KX_DEC_RESIDUAL && e_b < C.batch) {
                e_x[rd] = __ldcg(C.x + static_cast<long long>(e_b) * C.d_model + e_n);     // last written two phases ago
            }
        }
    }

    float acc[NT][4];
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) acc[tt][0] = acc[tt][1] = acc[tt][2] = acc[tt][3] = 0.f;
    float s1 = 0.f, s2 = 0.f;

    auto load_a = [&](int kc, uint4 (&av)[4]) {
        const int steps = min(CHUNK_K, p.K - kc) >> 5;
        const int spw = (steps + CONSUMER_WARPS - 1) / CONSUMER_WARPS;
        const int s_lo = warp * spw, s_hi = min(steps, s_lo + spw);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            av[u] = (aok && s_lo + u < s_hi) ? ld_cg4(ap + ((kc >> 3) + (s_lo + u) * 4)) : make_uint4(0, 0, 0, 0);
    };
    uint4 av[4];
    load_a(0, av);
    stamp(dbg, 1);
    for (int kc = 0; kc < p.K; kc += CHUNK_K) {
        const bool more = kc + CHUNK_K < p.K;
        const int steps = min(CHUNK_K, p.K - kc) >> 5;
        const int spw = (steps + CONSUMER_WARPS - 1) / CONSUMER_WARPS;        // <= 4
        const int s_lo = warp * spw, s_hi = min(steps, s_lo + spw);
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) {
            if (!tile_ok[tt]) continue;
            ring_wait(r.full + r.slot(), r.parity());
            if (kc == 0 && tt == 0) stamp(dbg, 2);
            const uint32_t wrow = r.addr() + g * W_PITCH + t * 16;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s_lo + u < s_hi) {
                    const uint4 wa = lds4(wrow + (s_lo + u) * 64);
                    const uint4 wb = lds4(wrow + 8 * W_PITCH + (s_lo + u) * 64);
                    mma16816(acc[tt], wa.x, wb.x, wa.y, wb.y, av[u].x, av[u].y);
                    mma16816(acc[tt], wa.z, wb.z, wa.w, wb.w, av[u].z, av[u].w);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(r.empty + r.slot());
            ++r.i;
        }
        if (ln) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (s_lo + u < s_hi) { stats2(av[u].x, s1, s2); stats2(av[u].y, s1, s2); stats2(av[u].z, s1, s2); stats2(av[u].w, s1, s2); }
        }
        // next chunk's activations: requested as soon as these registers are free (a second register set was measured:
        // no gain)
        if (more) load_a(kc + CHUNK_K, av);
    }
    stamp(dbg, 3);
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) {
        sm.red[tt][warp][g][2 * t] = acc[tt][0];
        sm.red[tt][warp][g][2 * t + 1] = acc[tt][1];
        sm.red[tt][warp][g + 8][2 * t] = acc[tt][2];
        sm.red[tt][warp][g + 8][2 * t + 1] = acc[tt][3];
    }
    if (ln) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        if (t == 0) { sm.st[warp][g][0] = s1; sm.st[warp][g][1] = s2; }
    }
    consumer_sync();
    stamp(dbg, 4);

    // per round, threads 0..127 finish tile 2*rd, threads 128..255 tile 2*rd + 1: (batch row b, feature r16)
    const int idx = tid & 127;
    const int b = idx >> 4, r16 = idx & 15;
    float vv[ROUNDS];
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
        const int tt = 2 * rd + (tid >> 7);
        const bool mine = tt < NT && item + tt * p.items < tiles;
        float v = 0.f;
        if (mine) {
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sm.red[tt][w][r16][b];
            if (ln) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { a1 += sm.st[w][b][0]; a2 += sm.st[w][b][1]; }
                const float inv_n = 1.0f / static_cast<float>(p.K);
                const float mean = a1 * inv_n;
                const float var = fmaxf(a2 * inv_n - mean * mean, 0.f);
                const float rstd = rsqrtf(var + C.eps);
                v = fmaf(-mean * rstd, e_lnc[rd], v * rstd);
            }
            v += e_bias[rd];
            sm.fin[tt][r16][b] = v;
        }
        vv[rd] = v;
    }
    consumer_sync();
    stamp(dbg, 5);

#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
        const int tt = 2 * rd + (tid >> 7);
        const bool mine = tt < NT && item + tt * p.items < tiles;
        const int n = (item + tt * p.items) * 16 + r16;
        const bool live = mine && b < C.batch && n < p.N;
        float v = vv[rd];
        if (p.argmax_keys != nullptr) {                     // all 32 lanes take part in the shuffles
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
            if (mine && r16 == 0 && b < C.batch && key != 0ull) atomicMax(p.argmax_keys + b, key);
        }
        if (live) {
            if (p.mode == KX_DEC_QKV) {
                const int which = n / C.d_model;
                const int col = n - which * C.d_model;
                if (which < 2) {
                    const float x0 = sm.fin[tt][r16 & ~1][b], x1 = sm.fin[tt][r16 | 1][b];
                    v = (r16 & 1) ? fmaf(x1, e_c[rd], x0 * e_s[rd]) : fmaf(x0, e_c[rd], -(x1 * e_s[rd]));
                }
                if (which == 0) {
                    p.q_out[static_cast<long long>(b) * C.d_model + col] = __float2bfloat16_rn(v);
                } else if (pos0 < C.t_max) {
                    __nv_bfloat16* dst = (which == 1 ? p.k_cache : p.v_cache);
                    dst[((static_cast<long long>(b) * C.heads + (col >> 6)) * C.t_max + pos0) * 64 + (col & 63)] =
                        __float2bfloat16_rn(v);
                }
            } else if (p.mode == KX_DEC_RESIDUAL) {
                float* px = C.x + static_cast<long long>(b) * C.d_model + n;
                v += e_x[rd];
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
    if (p.mode == KX_DEC_QKV) {
        // tell the attention items of this layer that one more tile of their head is in memory: release on the atomic,
        // cumulative over the CTA's stores ordered before it by the bar.sync (no grid barrier between q|k|v and attention)
        consumer_sync();
        if (tid < NT) {
            const int tile = item + tid * p.items;
            if (tile < tiles) {
                const int head = ((tile * 16) % C.d_model) >> 6;
                asm volatile("red.release.gpu.global.add.u64 [%0], 1;\n" :: "l"(C.barrier + HEAD_CNT0 + head) : "memory");
            }
        }
    }
    stamp(dbg, 6);
    // sm.red / sm.st are rewritten by this CTA's next item only after its k loop, sm.fin only after a further
    // consumer_sync: the reads above are ordered before those writes by the barriers in between.
}

// One (batch, head) item of the attention phase (consumer warps): all 256-key chunks, online softmax per lane.
__device__ __forceinline__ void attn_item(const Phase& p, const StepCommon& C, int bh, int n_keys, unsigned long long epoch,
                                          Ring& r, StepSmem& sm) {
    const int H = C.heads;
    const int b = bh / H, h = bh - b * H;
    // q and the newest k / v row of this head come from the q|k|v phase of this layer: wait for its 12 tiles (all other
    // heads may still be in flight — there is no grid barrier in front of the attention phase)
    if (threadIdx.x == 0) {
        const unsigned long long want = (epoch * C.layers + p.layer + 1) * TILES_PER_HEAD;
        const unsigned long long* cnt = C.barrier + HEAD_CNT0 + h;
        const long long t0 = clock64();
        while (ld_relaxed_u64(cnt) < want) {
            if (clock64() - t0 > (3ll << 30)) {
                if (C.err_flag != nullptr) atomicOr(C.err_flag, 4);
                break;
            }
        }
        (void)ld_acquire_u64(cnt);
    }
    consumer_sync();
    const int n_chunks = (n_keys + ATTN_CHUNK - 1) / ATTN_CHUNK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane & 7, kq = lane >> 3;
    const int newest = n_keys - 1;

    float qf[8];
    {
        const uint4 raw = ld_cg4(reinterpret_cast<const uint4*>(p.q_out + static_cast<long long>(b) * C.d_model + h * 64) + sub);
        const uint32_t rr[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            qf[2 * i] = __uint_as_float(rr[i] << 16) * C.scale_log2;
            qf[2 * i + 1] = __uint_as_float(rr[i] & 0xffff0000u) * C.scale_log2;
        }
    }
    // the newest key / value row is written by this launch's q|k|v phase: read it from L2 (one lane group owns it)
    const long long new_off = (static_cast<long long>(bh) * C.t_max + newest) * 64;
    const int lane_key = warp * 32 + kq;                     // key of iteration 0 within a chunk
    const int rel = (newest & (ATTN_CHUNK - 1)) - lane_key;  // newest == chunk*256 + lane_key + it*4  <=>  it = rel / 4
    const bool has_new = rel >= 0 && rel < 32 && (rel & 3) == 0;
    uint4 knew = make_uint4(0, 0, 0, 0);
    if (has_new) knew = ld_cg4(reinterpret_cast<const uint4*>(p.k_cache + new_off) + sub);
    float m = -INFINITY, l = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < n_chunks; ++c) {
        const int key0 = c * ATTN_CHUNK + lane_key;
        float sc[8];
        float mc = m;
        ring_wait(r.full + r.slot(), r.parity());
        {
            const uint32_t kbase = r.addr() + lane_key * 128 + sub * 16;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int j = key0 + it * 4;
                uint4 kr = make_uint4(0, 0, 0, 0);
                if (j < newest) kr = lds4(kbase + it * 512);
                else if (j == newest) kr = knew;
                const uint32_t rr[4] = {kr.x, kr.y, kr.z, kr.w};
                float d = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    d = fmaf(qf[2 * i], __uint_as_float(rr[i] << 16), d);
                    d = fmaf(qf[2 * i + 1], __uint_as_float(rr[i] & 0xffff0000u), d);
                }
                d += __shfl_xor_sync(0xffffffffu, d, 1);
                d += __shfl_xor_sync(0xffffffffu, d, 2);
                d += __shfl_xor_sync(0xffffffffu, d, 4);
                sc[it] = (j < n_keys) ? d : -INFINITY;
                mc = fmaxf(mc, sc[it]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.empty + r.slot());
        ++r.i;
        if (mc > m) {                                        // online softmax: rescale what this lane has so far
            const float f = m > -INFINITY ? ex2_approx(m - mc) : 0.f;
            l *= f;
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] *= f;
            m = mc;
        }
        uint4 vnew = make_uint4(0, 0, 0, 0);
        if (has_new && c == n_chunks - 1) vnew = ld_cg4(reinterpret_cast<const uint4*>(p.v_cache + new_off) + sub);
        ring_wait(r.full + r.slot(), r.parity());
        {
            const uint32_t vbase = r.addr() + lane_key * 128 + sub * 16;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int j = key0 + it * 4;
                uint4 vr = make_uint4(0, 0, 0, 0);
                if (j < newest) vr = lds4(vbase + it * 512);
                else if (j == newest) vr = vnew;
                const float pj = (j < n_keys) ? ex2_approx(sc[it] - m) : 0.f;
                l += pj;
                const uint32_t rr[4] = {vr.x, vr.y, vr.z, vr.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    o[2 * i] = fmaf(pj, __uint_as_float(rr[i] << 16), o[2 * i]);
                    o[2 * i + 1] = fmaf(pj, __uint_as_float(rr[i] & 0xffff0000u), o[2 * i + 1]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.empty + r.slot());
        ++r.i;
    }
    // merge the four key groups of the warp (different keys, different running maxima)
#pragma unroll
    for (int sft = 8; sft <= 16; sft <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, sft);
        const float lo = __shfl_xor_sync(0xffffffffu, l, sft);
        const float M = fmaxf(m, mo);
        const float f = m > -INFINITY ? ex2_approx(m - M) : 0.f;
        const float fo = mo > -INFINITY ? ex2_approx(mo - M) : 0.f;
        l = l * f + lo * fo;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float oo = __shfl_xor_sync(0xffffffffu, o[i], sft);
            o[i] = o[i] * f + oo * fo;
        }
        m = M;
    }
    consumer_sync();                                         // previous item's readers of att_o / att_ml are done
    if (kq == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.att_o[warp][sub * 8 + i] = o[i];
        if (sub == 0) { sm.att_ml[warp][0] = m; sm.att_ml[warp][1] = l; }
    }
    consumer_sync();
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
    const float4* p1 = C.text_off >= 0 ? reinterpret_cast<const float4*>(C.pos_table + static_cast<long long>(max(pr - C.text_off, 0)) * C.d_model) : nullptr;
    for (int i = threadIdx.x; i < (C.d_model >> 2); i += CONSUMERS) {
        float4 a = __ldg(e + i);
        const float4 c = __ldg(pp + i);
        if (p1 != nullptr) { const float4 c1 = __ldg(p1 + i); a.x += c1.x; a.y += c1.y; a.z += c1.z; a.w += c1.w; }
        const float4 rr = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
        reinterpret_cast<float4*>(C.x + static_cast<long long>(b) * C.d_model)[i] = rr;
        uint2 pk;
        pk.x = pack_bf16(rr.x, rr.y); pk.y = pack_bf16(rr.z, rr.w);
        reinterpret_cast<uint2*>(C.xb + static_cast<long long>(b) * C.d_model)[i] = pk;
    }
}

__device__ __forceinline__ void pick_tokens(const StepCommon& C) {
    const int step = ld_cg_i(C.step);
    const int pos = ld_cg_i(C.pos);
    consumer_sync();
    for (int b = threadIdx.x; b < C.batch; b += CONSUMERS) {
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(C.argmax_keys + b);
        long long choice = static_cast<long long>(0xffffffffu - static_cast<uint32_t>(key & 0xffffffffull));
        C.argmax_keys[b] = 0ull;
        if (C.forced != nullptr && step < C.hist_ld) choice = C.forced[static_cast<long long>(b) * C.hist_ld + step];
        C.tokens[b] = choice;
        if (C.history != nullptr && step < C.hist_ld) C.history[static_cast<long long>(b) * C.hist_ld + step] = choice;
    }
    consumer_sync();
    if (threadIdx.x == 0) {
        *C.step = step + 1;
        *C.pos = pos + 1;
    }
}

}  // namespace

__global__ void __maxnreg__(160)                         // one CTA of 9 warps per SM (the 6-stage ring fills shared memory); 3 warps share a 16 K-register sub-partition
decode_step_kernel(const StepPlan* __restrict__ plan) {
    extern __shared__ __align__(128) unsigned char ring_mem[];
    __shared__ StepSmem sm;
    static_assert(sizeof(StepCommon) % 4 == 0 && sizeof(Phase) % 4 == 0, "plan structs are copied as 32-bit words");
    stage_words(&sm.c, &plan->c, sizeof(StepCommon), threadIdx.x, STEP_THREADS);
    stage_words(&sm.ph[0], &plan->ph[0], sizeof(Phase), threadIdx.x, STEP_THREADS);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) { mbar_init(sm.full + s, 1); mbar_init(sm.empty + s, CONSUMER_WARPS); }
        fence_mbar_init();
    }
    __syncthreads();
    const StepCommon& C = sm.c;
    Ring r;
    r.base = smem_u32(ring_mem); r.full = sm.full; r.empty = sm.empty; r.i = 0;
    if (threadIdx.x >= CONSUMERS) {                          // producer warp: runs ahead through the whole step
        producer(plan, C, r);
        return;
    }
    const int n_phases = C.n_phases;
    unsigned long long* bar = C.barrier;
    const unsigned long long epoch = ld_acquire_u64(bar + 1);
    unsigned long long target = epoch * static_cast<unsigned long long>(n_phases - 1 - C.layers) * gridDim.x;
    const int pos0 = ld_cg_i(C.pos);                       // only the last phase of a launch changes it
    const int n_keys = min(pos0 + 1, C.t_max);

    for (int ph = 0; ph < n_phases; ++ph) {
        const Phase& P = sm.ph[ph & 1];
        const int type = P.type, items = P.items;
        long long* dbg = nullptr;
#ifdef KX_STEP_FINE_TRACE
        if (C.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && C.trace[2 * n_phases] == ph) dbg = C.trace + 2 * n_phases + 1;
#endif
        stamp(dbg, 11);
        if (type == PH_LINEAR) {
            // (one shared instantiation for both item kinds was measured too, to shrink the 56 KB kernel towards the 32 KB
            // L1.5 instruction cache: no gain, the one-tile phases just ran the longer code)
            if (P.nt > 2) {                                 // 3 or 4 tiles per item (the fourth predicated off when nt == 3)
                for (int item = blockIdx.x; item < items; item += gridDim.x) linear_item<4>(P, C, item, pos0, r, sm, dbg);
            } else if (P.nt == 2) {
                for (int item = blockIdx.x; item < items; item += gridDim.x) linear_item<2>(P, C, item, pos0, r, sm, dbg);
            } else {
                for (int item = blockIdx.x; item < items; item += gridDim.x) linear_item<1>(P, C, item, pos0, r, sm, dbg);
            }
        } else if (type == PH_ATTN) {
            for (int item = blockIdx.x; item < items; item += gridDim.x) attn_item(P, C, item, n_keys, epoch, r, sm);
        } else if (type == PH_EMBED) {
            if (static_cast<int>(blockIdx.x) < C.batch) embed_row(C, blockIdx.x);
        } else {
            if (blockIdx.x == 0) pick_tokens(C);
        }
        if (C.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) C.trace[2 * ph] = static_cast<long long>(globaltimer_ns());
        if (ph + 1 < n_phases && P.no_barrier) {            // q|k|v -> attention: per-head counters instead of a grid barrier
            consumer_sync();                                // everybody is done reading sm.ph[(ph + 1) & 1] (phase ph - 1)
            stage_words(&sm.ph[(ph + 1) & 1], &plan->ph[ph + 1], sizeof(Phase), threadIdx.x, CONSUMERS);
            consumer_sync();
            if (C.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) C.trace[2 * ph + 1] = static_cast<long long>(globaltimer_ns());
        } else if (ph + 1 < n_phases) {
            // the next phase's descriptor travels while the barrier drains (its buffer was last read in phase ph - 1)
            stage_words(&sm.ph[(ph + 1) & 1], &plan->ph[ph + 1], sizeof(Phase), threadIdx.x, CONSUMERS);
            target += gridDim.x;
            grid_barrier(bar, target, C.err_flag, dbg);
            stamp(dbg, 10);
            if (C.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) C.trace[2 * ph + 1] = static_cast<long long>(globaltimer_ns());
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence();
        bar[1] = epoch + 1;
    }
}

namespace {

constexpr size_t STEP_DYN_SMEM = static_cast<size_t>(NS) * STAGE_BYTES;

int step_grid(int sms) {
    static int per_sm = -1;
    if (per_sm < 0) {
        int n = 0;
        cudaError_t e = cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(STEP_DYN_SMEM));
        // two CTAs of 114 KB only fit with the full 228 KB carve-out; the default heuristic sized it for ONE block (the ncu
        // launch table showed a 148-CTA grid: half the intended parallelism and prefetch depth)
        if (e == cudaSuccess) e = cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, decode_step_kernel, STEP_THREADS, STEP_DYN_SMEM);
        if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
        per_sm = std::min(n, STEP_CTAS_PER_SM);
    }
    return per_sm * sms;
}

}  // namespace

}  // namespace kx

using namespace kx;

static int step_phase_count(int layers) { return 1 + 5 * layers + 2; }

extern "C" size_t kx_decode_plan_bytes(int layers) {
    if (layers < 0) return 0;
    return sizeof(StepPlan) + sizeof(Phase) * static_cast<size_t>(step_phase_count(layers));
}

extern "C" int kx_decode_plan_build(const kx_decode_step_args* g, void* device_plan, cudaStream_t stream) {
    if (!g || !device_plan) { set_error("kx_decode_plan_build: null argument"); return KX_ERR_ARG; }
    if (g->batch <= 0 || g->batch > 8 || g->layers < 0 || g->d_model <= 0 || (g->d_model & 63) || g->ffn <= 0 || (g->ffn & 31) ||
        g->heads * 64 != g->d_model || g->heads > 256 || g->vocab <= 0 || g->t_max <= 0 || g->pos_rows <= 2) {
        set_error("kx_decode_plan_build: need 1 <= batch <= 8, d_model == heads*64, ffn %% 32 == 0 (batch %d, d_model %d, heads %d)",
                  g->batch, g->d_model, g->heads);
        return KX_ERR_ARG;
    }
    const void* need[] = {g->w_qkv, g->c_qkv, g->d_qkv, g->w_o, g->c_o, g->d_o, g->w_fc1, g->c_fc1, g->d_fc1, g->w_fc2, g->c_fc2,
                          g->d_fc2, g->k_cache, g->v_cache, g->w_out, g->c_out, g->embed_table, g->pos_table, g->xq_cos, g->xq_sin,
                          g->xk_cos, g->xk_sin, g->tokens, g->x, g->xb, g->q, g->att, g->mid, g->logits, g->argmax_keys, g->pos,
                          g->step, g->barrier};
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
    C.batch = g->batch; C.d_model = D; C.heads = g->heads; C.t_max = g->t_max; C.vocab = g->vocab; C.pos_rows = g->pos_rows;
    C.text_off = g->text_index_off;
    C.hist_ld = g->history_ld; C.n_phases = n_ph;
    C.eps = g->eps; C.scale_log2 = g->scale * 1.4426950408889634f;
    C.forced = g->forced; C.tokens = g->tokens; C.history = g->history;
    C.embed_table = g->embed_table; C.pos_table = g->pos_table;
    C.xq_cos = g->xq_cos; C.xq_sin = g->xq_sin; C.xk_cos = g->xk_cos; C.xk_sin = g->xk_sin;
    C.x = g->x; C.xb = reinterpret_cast<__nv_bfloat16*>(g->xb);
    C.pos = g->pos; C.step = g->step; C.err_flag = g->err_flag; C.argmax_keys = g->argmax_keys;
    C.barrier = g->barrier;
    C.layers = L;
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
        (void)allow_split;
        P.nt = std::min(4, (tiles + grid - 1) / grid);      // more tiles than CTAs: group them so the phase stays one pass
        P.items = (tiles + P.nt - 1) / P.nt;
        return P;
    };
    plan->ph[k].type = PH_EMBED; plan->ph[k].items = g->batch; ++k;
    for (int l = 0; l < L; ++l) {
        Phase& q = lin(g->xb, D, g->w_qkv[l], 3 * D, D, g->c_qkv[l], g->d_qkv[l], KX_DEC_QKV, KX_ACT_NONE, false);
        q.q_out = reinterpret_cast<__nv_bfloat16*>(g->q);
        q.k_cache = reinterpret_cast<__nv_bfloat16*>(g->k_cache[l]); q.v_cache = reinterpret_cast<__nv_bfloat16*>(g->v_cache[l]);
        q.layer = l; q.no_barrier = 1;
        Phase& at = plan->ph[k++];
        at.type = PH_ATTN; at.items = g->batch * g->heads; at.layer = l;
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
        if (plan->ph[i].type == PH_LINEAR && (plan->ph[i].K & 31)) {
            set_error("kx_decode_plan_build: unsupported Linear shape (K %d)", plan->ph[i].K);
            return KX_ERR_ARG;
        }
    // one-time setup call: the plan is copied from a temporary host buffer, so the stream is synchronised here
    cudaError_t e = cudaMemcpyAsync(device_plan, host.data(), bytes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("kx_decode_plan_build: %s", cudaGetErrorString(e)); return KX_ERR_LAUNCH; }
    return KX_OK;
}

extern "C" int kx_decode_step_ctas(void) {
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    return step_grid(sms);
}

extern "C" int kx_decode_step(const void* device_plan, cudaStream_t stream) {
    if (!device_plan) { set_error("kx_decode_step: null plan"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const int grid = step_grid(sms);
    if (grid <= 0) { set_error("kx_decode_step: decode_step_kernel does not fit on an SM"); return KX_ERR_LAUNCH; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(STEP_THREADS); cfg.dynamicSmemBytes = STEP_DYN_SMEM; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;              // co-residency of the whole grid is what the barrier relies on
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const StepPlan* plan = reinterpret_cast<const StepPlan*>(device_plan);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, decode_step_kernel, plan);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("kx_decode_step launch failed: %s", cudaGetErrorString(e));
        return KX_ERR_LAUNCH;
    }
    count_launch();
    return KX_OK;
}
