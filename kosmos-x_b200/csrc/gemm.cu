// kx_gemm_bf16: C[M,N] = epilogue(A[M,K] . W[N,K]^T)   — the Linear layers of the Kosmos-X path.
//
// Replaces every F.linear / addmm the reference reaches through torchscale / HF CLIP /
// flamingo_pytorch (SURVEY.md §2.4 k1,k3,k5,k6,k7,k11,k15,k16,k17,k18).
//
// sm_100a design: persistent, warp-specialised.  One CTA (CG=1) or one CTA pair (CG=2,
// cta_group::2) per SM / SM pair owns a (128*CG) x BN output tile per iteration:
//   warp 0   TMA producer   : cp.async.bulk.tensor (SWIZZLE_128B) -> smem ring (mbarrier full/empty)
//   warp 1   MMA issuer     : tcgen05.mma kind::f16 (bf16 x bf16 -> fp32) into TMEM, 2 accumulator
//                             stages of BN columns; tcgen05.commit releases smem slots / signals epilogue
//   warp 2   TMEM allocator
//   warps 4-7 epilogue      : tcgen05.ld 32x32b -> registers -> fused bias / xPos rotation / GELU /
//                             residual / positional add / row scatter -> global (bf16 or fp32)
// The epilogue of tile i overlaps the MMAs of tile i+1 (double-buffered TMEM accumulator).
#include "kx_internal.h"
#include "ptx.cuh"
#include "philox.cuh"

#include <mutex>
#include <unordered_map>

namespace kx {

// Build with -DKX_GEMM_TRACE to record clock64 stamps of epilogue warp 4 of CTA 0 (staged epilogue):
// trace[(tile*8 + chunk)*8 + point]; kx_gemm_set_trace installs the device buffer.  Off in the shipped library.
#ifdef KX_GEMM_TRACE
__device__ long long* g_gemm_trace = nullptr;
#define KX_GT(tile, chunk, point)                                                                       \
    do {                                                                                                \
        if (g_gemm_trace != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0 && (tile) < 16)        \
            g_gemm_trace[(((tile)*8 + (chunk)) * 8) + (point)] = clock64();                             \
    } while (0)
#else
#define KX_GT(tile, chunk, point) do {} while (0)
#endif

constexpr int BLOCK_M = 128;   // rows per CTA (UMMA M = 128 * CG)
constexpr int BLOCK_K = 64;    // 64 bf16 = one 128-byte swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // warps 0-3: TMA producer, MMA issuer, TMEM allocator, idle; warps 4-11: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;       // two per SM sub-partition: warp e handles TMEM lanes 32*(e&3).. and column half e>>2

// TMA_EPI: the epilogue stages each warp's 32-row x 128-byte sub-tile in swizzled shared memory and
// writes it with cp.async.bulk.tensor (coalesced, asynchronous); the fp32 residual sub-tile is
// prefetched the same way.  Costs 16 KB of staging per epilogue warp, taken from the operand ring.
constexpr int EPI_STAGE_BYTES = 32 * 128;                    // 32 rows x 128 B
constexpr int EPI_WARP_BYTES = 2 * EPI_STAGE_BYTES;          // main 4 KB + copy[2] x 2 KB

template <int CG, int BN, bool TMA_EPI>
struct GemmCfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_ROWS = BN / CG;
    static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_BYTES = TMA_EPI ? EPI_WARPS * EPI_WARP_BYTES : 0;
    static constexpr int RING_BUDGET = TMA_EPI ? (160 * 1024) : (192 * 1024);
    static constexpr int STAGES = RING_BUDGET / STAGE_BYTES;
    static constexpr int VEC_BYTES = TMA_EPI ? 2 * BN * 4 : 0;   // this tile's bias[] and ln_c[] columns, staged once per tile
    static constexpr int BAR_BYTES = 512;                       // (2*STAGES + 4 + 16) mbarriers + the TMEM slot; STAGES <= 8
    static_assert((2 * STAGES + 20) * 8 + 8 <= BAR_BYTES, "barrier block");
    static_assert(STAGES >= 2, "ring too shallow");
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + VEC_BYTES + BAR_BYTES;   // base must be 1 KB aligned
    static constexpr int TMEM_COLS = 2 * BN;                                    // power of two for BN in {64,128,256}
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct GemmEpi {
    int M, N, K;
    int raster_g, raster_nc, raster_nfast;   // tile order (tile_coords): m-blocks per group, n-blocks per chunk, which index runs fastest
    const float* bias;         // [N] fp32 or null
    const float* res;          // fp32 residual, indexed like out (row remap applied), or null
    long long ld_res;
    void* out;                 // bf16 or fp32
    long long ld_out;
    int act;                   // KX_ACT_*
    int grp_rows, grp_stride, grp_off;   // out_row = (m / grp_rows) * grp_stride + grp_off + m % grp_rows
    const float* add_tab;      // fp32 [*, ld_add]; row (m % grp_rows) + add_off is added (positional tables)
    int add_off;
    long long ld_add;
    const float *xq_cos, *xq_sin, *xk_cos, *xk_sin;   // xPos tables [seq_len, 32] fp32
    int seq_len, d_model;
    int vec_ok;                // 16-byte vector access to out/res/add_tab rows is legal
    int tma_store;             // staged epilogue: 1 = TMA store, 0 = coalesced float2 stores from the staging buffer
    // LayerNorm folded into this GEMM: A holds the RAW rows, W already carries gamma,
    //   y = rstd[m] * (acc - mean[m] * ln_c[n]) + bias[n],   bias = W.beta + b
    // mean/rstd come from ln_tiles partial (sum, sumsq) pairs per row written by the producer's epilogue.
    const float2* ln_part;     // [ln_tiles][M] or null
    const float* ln_c;         // [N]: sum_k W'[n,k]
    int ln_tiles;
    float ln_inv_n, ln_eps;
    // producer side: per-row partial (sum, sumsq) of the values this GEMM stores (after bf16 rounding),
    // one pair per (128-column block, row); and an optional bf16 copy of an fp32 output (staged epilogue only)
    float2* stats_out;         // [ceil(N/(BN/2))][M] or null
    void* out2;                // bf16 [M, ld_out2] or null
    long long ld_out2;
    // training: dropout on the Linear's output (after bias / activation, BEFORE the residual add) — torchscale's
    // `x = dropout(self_attn(...))`, `x = dropout(fc2(...))`; element (m, n) is kept iff drop_keep8(drop, m, n / 8) bit n % 8
    DropSpec drop;             // thr == 0: off
};

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int G, int NC, bool n_fast, int& m_blk, int& n_blk) {
    // Raster: groups of G m-blocks; inside a group, chunks of up to NC n-blocks; the tiles in flight at one time (one per
    // cluster) then cover ~G x NC blocks, A and B footprints balanced.  Inside a chunk the index that runs fastest decides
    // which operand's sharers get ADJACENT cluster ids (cluster launches map block ids to SM ids contiguously): n_fast puts
    // the clusters sharing an A row-block next to each other.  Measured (profiles/r2_gemm_raster.md): n fastest in 8 x 8
    // chunks cuts fc2's DRAM reads 915 -> 750 MB per launch; G = 16 when the weights are the larger operand (LM head) halves
    // the passes over W, 1.19 -> 0.93 GB; step time equal within noise (the step is power-bound).  kx_gemm_bf16 picks them.
    const int per_group = G * num_n;
    const int g = t / per_group;
    const int first_m = g * G;
    const int gsz = min(G, num_m - first_m);
    const int r = t - g * per_group;
    const int chunk = r / (gsz * NC);
    const int w = min(NC, num_n - chunk * NC);
    const int rr = r - chunk * gsz * NC;
    m_blk = first_m + (n_fast ? rr / w : rr % gsz);
    n_blk = chunk * NC + (n_fast ? rr % w : rr / gsz);
}

// bias -> [xPos rotation] -> activation on one 32-column chunk of row m (shared by both store paths)
// (sum, sumsq) of eight stored bf16 values, accumulated as packed f32x2 pairs on two independent chains
// (acc[0..1] sums, acc[2..3] sums of squares; folded by stats_fold).  `full` = no N tail in this chunk.
__device__ __forceinline__ void stats_add8(const uint4& pk, bool full, int n, int N, uint64_t (&acc)[4]) {
    const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float lo = __uint_as_float(w[u] << 16), hi = __uint_as_float(w[u] & 0xffff0000u);
        if (!full) {
            if (n + 2 * u >= N) lo = 0.f;
            if (n + 2 * u + 1 >= N) hi = 0.f;
        }
        const uint64_t v = pack_f32x2(lo, hi);
        acc[u & 1] = fadd2(acc[u & 1], v);
        acc[2 + (u & 1)] = ffma2(v, v, acc[2 + (u & 1)]);
    }
}
__device__ __forceinline__ void stats_fold(const uint64_t (&acc)[4], float& s1, float& s2) {
    float a0, a1, b0, b1;
    unpack_f32x2(fadd2(acc[0], acc[1]), a0, a1);
    unpack_f32x2(fadd2(acc[2], acc[3]), b0, b1);
    s1 = a0 + a1;
    s2 = b0 + b1;
}

// Row statistics of the A operand for a folded LayerNorm: fixed-order sum of the producer's partials.
__device__ __forceinline__ float2 ln_row_stats(const GemmEpi& ep, int m) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < ep.ln_tiles; ++i) {
        const float2 p = __ldg(ep.ln_part + static_cast<long long>(i) * ep.M + m);
        s1 += p.x; s2 += p.y;
    }
    const float mean = s1 * ep.ln_inv_n;
    const float var = fmaxf(s2 * ep.ln_inv_n - mean * mean, 0.f);
    return make_float2(mean, rsqrtf(var + ep.ln_eps));
}

// sb / sc: this chunk's 32 bias / ln_c values staged in shared memory (zero padded past N), or null = read global.
template <int EPI>
__device__ __forceinline__ void epilogue_math(const GemmEpi& ep, float (&f)[32], int m, int n0, bool full, float2 ln,
                                              const float* sb = nullptr, const float* sc = nullptr) {
    if (sb != nullptr) {
        if (ep.ln_part != nullptr) {
            const float nmr = -ln.x * ln.y;             // y = rstd*acc + (-mean*rstd)*c
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 c = *reinterpret_cast<const float4*>(sc + i);
                f[i] = fmaf(nmr, c.x, f[i] * ln.y); f[i + 1] = fmaf(nmr, c.y, f[i + 1] * ln.y);
                f[i + 2] = fmaf(nmr, c.z, f[i + 2] * ln.y); f[i + 3] = fmaf(nmr, c.w, f[i + 3] * ln.y);
            }
        }
        if (ep.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b = *reinterpret_cast<const float4*>(sb + i);
                f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
        }
    } else {
    if (ep.ln_part != nullptr) {
        const float nmr = -ln.x * ln.y;                 // y = rstd*acc + (-mean*rstd)*c
        if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(ep.ln_c + n0 + i));
                f[i] = fmaf(nmr, c.x, f[i] * ln.y); f[i + 1] = fmaf(nmr, c.y, f[i + 1] * ln.y);
                f[i + 2] = fmaf(nmr, c.z, f[i + 2] * ln.y); f[i + 3] = fmaf(nmr, c.w, f[i + 3] * ln.y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < ep.N) f[i] = fmaf(nmr, __ldg(ep.ln_c + n0 + i), f[i] * ln.y);
        }
    }
    if (ep.bias != nullptr) {
        if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + i));
                f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < ep.N) f[i] += __ldg(ep.bias + n0 + i);
        }
    }
    }

    if constexpr (EPI == KX_EPI_QKV_XPOS) {
        // columns [0,d) = q (upscale tables), [d,2d) = k (downscale tables), [2d,3d) = v (untouched).
        // Pair j of a 64-wide head = columns (2j, 2j+1): out0 = x0*c - x1*s ; out1 = x1*c + x0*s.
        const int which = n0 / ep.d_model;
        if (which < 2) {
            const int t = m % ep.seq_len;
            const int j0 = (n0 & 63) >> 1;
            const float* ct = (which == 0 ? ep.xq_cos : ep.xk_cos) + t * 32 + j0;
            const float* st = (which == 0 ? ep.xq_sin : ep.xk_sin) + t * 32 + j0;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(ct + i));
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(st + i));
                const float c[4] = {c4.x, c4.y, c4.z, c4.w};
                const float s[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float x0 = f[2 * (i + u)], x1 = f[2 * (i + u) + 1];
                    f[2 * (i + u)] = x0 * c[u] - x1 * s[u];
                    f[2 * (i + u) + 1] = x1 * c[u] + x0 * s[u];
                }
            }
        }
    }

    if (ep.act == KX_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) gelu_erf_x2(f[i], f[i + 1]);
    } else if (ep.act == KX_ACT_QUICK_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = quick_gelu(f[i]);
    }
    if (ep.drop.thr != 0u) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t keep = drop_keep8(ep.drop, static_cast<uint32_t>(m), static_cast<uint32_t>((n0 >> 3) + i));
#pragma unroll
            for (int u = 0; u < 8; ++u) f[8 * i + u] = ((keep >> u) & 1u) ? f[8 * i + u] * ep.drop.inv_keep : 0.f;
        }
    }
}

template <bool OUT_F32, int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& ep, const uint32_t (&v)[32], int m, int n0, float2 ln) {
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
    const bool full = (n0 + 32 <= ep.N);
    epilogue_math<EPI>(ep, f, m, n0, full, ln);

    long long orow = m;
    int prow = 0;
    if (ep.grp_rows > 0) {
        const int g = m / ep.grp_rows;
        prow = m - g * ep.grp_rows;
        orow = static_cast<long long>(g) * ep.grp_stride + ep.grp_off + prow;
    }
    const bool vec = full && ep.vec_ok;

    if (ep.add_tab != nullptr) {
        const float* a = ep.add_tab + static_cast<long long>(prow + ep.add_off) * ep.ld_add + n0;
        if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(a + i));
                f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < ep.N) f[i] += __ldg(a + i);
        }
    }
    if (ep.res != nullptr) {
        const float* r = ep.res + orow * ep.ld_res + n0;
        if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b = *reinterpret_cast<const float4*>(r + i);
                f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < ep.N) f[i] += r[i];
        }
    }

    if constexpr (OUT_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + orow * ep.ld_out + n0;
        if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
        } else if (full && ((ep.ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(ep.out) & 7) == 0)) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) *reinterpret_cast<float2*>(o + i) = make_float2(f[i], f[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < ep.N) o[i] = f[i];
        }
    } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + orow * ep.ld_out + n0;
        if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                uint4 q;
                q.x = pack_bf16(f[i], f[i + 1]);
                q.y = pack_bf16(f[i + 2], f[i + 3]);
                q.z = pack_bf16(f[i + 4], f[i + 5]);
                q.w = pack_bf16(f[i + 6], f[i + 7]);
                *reinterpret_cast<uint4*>(o + i) = q;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + i < ep.N) o[i] = __float2bfloat16_rn(f[i]);
        }
    }
}

// ATR / BTR: the operand is given TRANSPOSED in memory — A as [K, M] row-major, W as [K, N] row-major — i.e. its
// MN dimension is the contiguous one.  The tile is then staged as 64-column atoms ([64 k-rows][128 B], SWIZZLE_128B,
// one TMA box each) and fed to tcgen05.mma as an MN-major operand (instruction-descriptor major bits; descriptor
// LBO = 8 KB between MN atoms, SBO = 1 KB between 8-row k groups; a UMMA_K step advances 16 rows = 2 KB).
// This is how the backward GEMMs read their operands where they lie: dgrad dX = dY . W (W as [K', N']) and
// wgrad dW = dY^T . X (both operands [K', *]) need no transposed copies.
constexpr int MN_ATOM_BYTES = 64 * 128;

template <int CG, int BN, bool OUT_F32, int EPI, bool TMA_EPI, bool ATR = false, bool BTR = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmOut2, const GemmEpi ep) {
    using Cfg = GemmCfg<CG, BN, TMA_EPI>;
    constexpr int STAGES = Cfg::STAGES;
    // SWIZZLE_128B atoms need 1024-byte alignment (identical offset in both CTAs of a pair); the kernel has no
    // static shared memory, so the dynamic window starts at the aligned base of the CTA's allocation.
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023)) { printf("kx gemm: dynamic smem base not 1024-aligned\n"); __trap(); }
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;          // 1024-aligned (stage sizes are multiples of 1 KB)
    float* s_vec = reinterpret_cast<float*>(smem_epi + Cfg::EPI_BYTES);    // [2][BN]: bias, ln_c of the current tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::EPI_BYTES + Cfg::VEC_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint64_t* resbar = bars + 2 * STAGES + 4;                      // [8 warps]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = (rank == 0);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if constexpr (TMA_EPI) {
            prefetch_tmap(&tmOut);
            if (ep.res != nullptr) prefetch_tmap(&tmRes);
            if (ep.out2 != nullptr) prefetch_tmap(&tmOut2);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], CG);       // one arrival per CTA producer, on the leader's barrier
            mbar_init(&empty[s], 1);       // one tcgen05.commit (multicast to both CTAs when CG=2)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], CG * EPI_WARPS * 32);   // every epilogue thread of every CTA, on the leader's barrier
        }
        for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&resbar[i], 1);
        fence_mbar_init();
    }
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == 2) {
        tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_m = (ep.M + BLOCK_M * CG - 1) / (BLOCK_M * CG);
    const int num_n = (ep.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (ep.K + BLOCK_K - 1) / BLOCK_K;
    const int cluster_id = blockIdx.x / CG;
    const int num_clusters = gridDim.x / CG;

    // Single-thread roles are entered through elect.sync: ptxas then knows exactly one lane is active and emits
    // the uniform-datapath instructions (UTMALDG / UTCHMMA) straight-line, without a per-instruction
    // divergence loop (measurably cheaper issue than `lane == 0`).
    if (warp < EPI_WARP0) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");      // registers go to the epilogue warpgroups
    if (warp == 0) {
        if (elect_one()) {
        // ================= TMA producer =================
        int s = 0;
        uint32_t ph = 0;
        // W (K-major B) is re-read by every m-block group; the activations stream through once per group: keep W in L2
        const uint64_t b_hint = (!ATR && !BTR) ? kEvictLast : kEvictNormal;
        for (int t = cluster_id; t < num_tiles; t += num_clusters) {
            int m_blk, n_blk;
            tile_coords(t, num_m, num_n, ep.raster_g, ep.raster_nc, ep.raster_nfast != 0, m_blk, n_blk);
            const int m0 = m_blk * (BLOCK_M * CG) + rank * BLOCK_M;
            const int n0 = n_blk * BN + rank * Cfg::B_ROWS;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* sa = smem_a + s * Cfg::A_BYTES;
                uint8_t* sb = smem_b + s * Cfg::B_BYTES;
                if constexpr (CG == 1) {
                    mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
                    if constexpr (!ATR) tma_load_2d(&tmA, &full[s], sa, kb * BLOCK_K, m0);
                    else
#pragma unroll
                        for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d(&tmA, &full[s], sa + c * MN_ATOM_BYTES, m0 + c * 64, kb * BLOCK_K);
                    if constexpr (!BTR) tma_load_2d(&tmB, &full[s], sb, kb * BLOCK_K, n0, b_hint);
                    else
#pragma unroll
                        for (int c = 0; c < Cfg::B_ROWS / 64; ++c) tma_load_2d(&tmB, &full[s], sb + c * MN_ATOM_BYTES, n0 + c * 64, kb * BLOCK_K);
                } else {
                    if (leader) mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES * 2);
                    else mbar_arrive_cluster(&full[s], 0);
                    if constexpr (!ATR) tma_load_2d_cg2(&tmA, &full[s], sa, kb * BLOCK_K, m0);
                    else
#pragma unroll
                        for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d_cg2(&tmA, &full[s], sa + c * MN_ATOM_BYTES, m0 + c * 64, kb * BLOCK_K);
                    if constexpr (!BTR) tma_load_2d_cg2(&tmB, &full[s], sb, kb * BLOCK_K, n0, b_hint);
                    else
#pragma unroll
                        for (int c = 0; c < Cfg::B_ROWS / 64; ++c) tma_load_2d_cg2(&tmB, &full[s], sb + c * MN_ATOM_BYTES, n0 + c * 64, kb * BLOCK_K);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
        // ================= MMA issuer (leader CTA only) =================
        constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M * CG, BN, ATR ? 1u : 0u, BTR ? 1u : 0u);
        constexpr uint32_t a_step = ATR ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;     // descriptor address units of 16 B
        constexpr uint32_t b_step = BTR ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t adesc = make_desc_sw128(smem_u32(smem_a + s * Cfg::A_BYTES), ATR ? MN_ATOM_BYTES : 0);
                const uint64_t bdesc = make_desc_sw128(smem_u32(smem_b + s * Cfg::B_BYTES), BTR ? MN_ATOM_BYTES : 0);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                    umma_bf16<CG>(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) != 0);
                if constexpr (CG == 1) umma_commit(&empty[s]); else umma_commit_cg2(&empty[s], 0b11);
                if (kb == num_kb - 1) {
                    if constexpr (CG == 1) umma_commit(&tfull[a]); else umma_commit_cg2(&tfull[a], 0b11);
                }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
        if constexpr (CG == 2) {
            // the peer's last remote arrivals must land before this CTA's barriers go away
            if (it > 0) {
                const int last = it - 1;
                mbar_wait(&tempty[last & 1], (last >> 1) & 1);
            }
        }
        }
    } else if (warp >= EPI_WARP0) {
        // ================= epilogue =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int ew = warp - EPI_WARP0;              // 0..7
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int chalf = ew >> 2;                    // which half of the tile's columns
        constexpr int CPW = BN / 64;                  // 32-column chunks per warp
        const int c_begin = chalf * CPW;
        int it = 0;
        if constexpr (!TMA_EPI) {
            for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
                int m_blk, n_blk;
                tile_coords(t, num_m, num_n, ep.raster_g, ep.raster_nc, ep.raster_nfast != 0, m_blk, n_blk);
                const int a = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                const int m = m_blk * (BLOCK_M * CG) + rank * BLOCK_M + q * 32 + lane;
                const int nb = n_blk * BN;
                mbar_wait(&tfull[a], aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
                const float2 ln = (ep.ln_part != nullptr && m < ep.M) ? ln_row_stats(ep, m) : make_float2(0.f, 1.f);
#pragma unroll 1
                for (int c = c_begin; c < c_begin + CPW; ++c) {
                    if (nb + c * 32 >= ep.N) break;       // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
                    if (m < ep.M) epilogue_chunk<OUT_F32, EPI>(ep, v, m, nb + c * 32, ln);
                }
                tc_fence_before();
                if constexpr (CG == 1) mbar_arrive(&tempty[a]); else mbar_arrive_cluster(&tempty[a], 0);
            }
        } else {
            // Staged epilogue.  Eight warps; warp (q, chalf) owns rows 32q..32q+31 and one column half of every tile.
            // Per warp, private smem: main 4 KB (32 rows x 128 B, SWIZZLE_128B) and copy[2] x 2 KB (32 rows x 64 B,
            // SWIZZLE_64B) for the bf16 copy of an fp32 output.
            //   fp32 out: unit = one 32-column chunk.  With a residual, `main` first receives the residual sub-tile
            //             by TMA (issued at unit start, landing during the TMEM load and the math) and is updated IN PLACE.
            //   bf16 out: unit = two chunks (64 columns = 128-byte rows).
            // (Round 2 measured two alternatives for the residual and kept this form: a dedicated third buffer per warp so that
            // the residual of chunk c+1 lands while chunk c is stored, paid for with one operand stage — out_proj 136.7 us against
            // 127.8 us here; and plain row loads into registers one chunk ahead — 254 us, the per-thread 128-byte row segments
            // cost 32 L1 wavefronts per load instruction.  profiles/r2_gemm_epilogue_experiments.md.)
            // One cp.async.bulk group per store.  `main` is single-buffered: before it is overwritten (by the residual
            // load at unit start, or by the first write of the unit, which comes after the math) lane 0 waits until the
            // previous unit's store has finished reading it; the second warp on the SM sub-partition covers that latency.
            uint8_t* st_main = smem_epi + ew * EPI_WARP_BYTES;
            uint8_t* st_copy = st_main + EPI_STAGE_BYTES;
            uint64_t* rbar = resbar + ew;
            const bool has_res = (ep.res != nullptr);
            const bool has_copy = OUT_F32 && (ep.out2 != nullptr);
            const bool has_stats = (ep.stats_out != nullptr);
            const bool tma_out = ep.tma_store != 0;
            const int sw = lane & 7;
            const int sw64 = (lane >> 1) & 3;
            uint32_t unit = 0;              // store units issued by this warp
            auto wait_main_free = [&]() {   // group order per unit: main store, then copy store
                if (lane == 0) { if (has_copy) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
                __syncwarp();
            };
            for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
                int m_blk, n_blk;
                tile_coords(t, num_m, num_n, ep.raster_g, ep.raster_nc, ep.raster_nfast != 0, m_blk, n_blk);
                const int a = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                const int m_row0 = m_blk * (BLOCK_M * CG) + rank * BLOCK_M + q * 32;
                const int m = m_row0 + lane;
                const int nb = n_blk * BN;
                const int nchunks = min(c_begin + CPW, (ep.N - nb + 31) >> 5);   // this warp walks chunks [c_begin, nchunks)
                const bool rows_ok = m_row0 < ep.M;          // warp-uniform; rows >= M inside a box are clipped by TMA
                const float2 ln = (ep.ln_part != nullptr && m < ep.M) ? ln_row_stats(ep, m) : make_float2(0.f, 1.f);
                float s1 = 0.f, s2 = 0.f;
                uint64_t sacc[4] = {0ull, 0ull, 0ull, 0ull};
                KX_GT(it, 0, 0);
                {   // stage this tile's bias / ln_c columns once (the 8 epilogue warps walk the same tile sequence)
                    const int col = ew * 32 + lane;                     // 256 threads, BN <= 256 columns
                    const bool in = col < BN && nb + col < ep.N;
                    const float bv = (in && ep.bias != nullptr) ? __ldg(ep.bias + nb + col) : 0.f;
                    const float cv = (in && ep.ln_part != nullptr) ? __ldg(ep.ln_c + nb + col) : 0.f;
                    named_bar_sync(1, EPI_WARPS * 32);                  // previous tile's readers are done
                    if (col < BN) { s_vec[col] = bv; s_vec[BN + col] = cv; }
                    named_bar_sync(1, EPI_WARPS * 32);
                }
                if (has_res && lane < CPW) {
                    // the residual comes from HBM (a 134 MB stream): pull the NEXT tile's sub-tiles of this warp into L2 now,
                    // a whole tile time ahead, so that the single-buffered TMA loads below only see L2 latency
                    const int tn = t + num_clusters;
                    if (tn < num_tiles) {
                        int mb2, nb2;
                        tile_coords(tn, num_m, num_n, ep.raster_g, ep.raster_nc, ep.raster_nfast != 0, mb2, nb2);
                        const int mr = mb2 * (BLOCK_M * CG) + rank * BLOCK_M + q * 32;
                        const int nc = nb2 * BN + (c_begin + lane) * 32;
                        if (mr < ep.M && nc < ep.N) tma_prefetch_l2_2d(&tmRes, nc, mr);
                    }
                }
                auto issue_res = [&](int c) {               // residual sub-tile of chunk c -> main (freed by wait_main_free)
                    if (lane == 0) {
                        mbar_arrive_expect_tx(rbar, EPI_STAGE_BYTES);
                        tma_load_2d(&tmRes, rbar, st_main, nb + c * 32, m_row0);
                    }
                };
                KX_GT(it, 0, 1);
                mbar_wait(&tfull[a], aph);
                tc_fence_after();
                KX_GT(it, 0, 2);
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN;
                uint32_t v[32];
                if (c_begin < nchunks) tmem_ld32(taddr + c_begin * 32, v);
#pragma unroll 1
                for (int c = c_begin; c < nchunks; ++c) {
                    const int n0 = nb + c * 32;
                    KX_GT(it, c, 3);
                    const bool unit_start = OUT_F32 || !(c & 1);
                    if (rows_ok && has_res) {                // residual: the buffer must be free before the load is issued
                        wait_main_free();
                        issue_res(c);
                    }
                    KX_GT(it, c, 4);
                    float f[32];
                    tmem_ld_wait();
                    KX_GT(it, c, 5);
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    if (c + 1 < nchunks) tmem_ld32(taddr + (c + 1) * 32, v);      // in flight during this chunk's math
                    if (!rows_ok) continue;
                    const bool full = n0 + 32 <= ep.N;
                    epilogue_math<EPI>(ep, f, m, n0, full, ln, s_vec + c * 32, s_vec + BN + c * 32);
                    KX_GT(it, c, 6);
                    uint8_t* mb = st_main + lane * 128;
                    if (unit_start && !has_res) wait_main_free();      // after the math: the previous store has had time to drain
                    if constexpr (OUT_F32) {
                        if (has_res) {
                            mbar_wait(rbar, unit & 1);
                            KX_GT(it, c, 7);
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                const float4 x = *reinterpret_cast<const float4*>(mb + ((g ^ sw) << 4));
                                f[4 * g] += x.x; f[4 * g + 1] += x.y; f[4 * g + 2] += x.z; f[4 * g + 3] += x.w;
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            *reinterpret_cast<float4*>(mb + ((g ^ sw) << 4)) = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
                        if (has_copy) {
                            uint8_t* cb = st_copy + (unit & 1) * (EPI_STAGE_BYTES / 2) + lane * 64;
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 pk;
                                pk.x = pack_bf16(f[8 * g + 0], f[8 * g + 1]);
                                pk.y = pack_bf16(f[8 * g + 2], f[8 * g + 3]);
                                pk.z = pack_bf16(f[8 * g + 4], f[8 * g + 5]);
                                pk.w = pack_bf16(f[8 * g + 6], f[8 * g + 7]);
                                *reinterpret_cast<uint4*>(cb + ((g ^ sw64) << 4)) = pk;
                                if (has_stats)            // statistics of what the consumer GEMM will read: the bf16 copy
                                    stats_add8(pk, full, n0 + 8 * g, ep.N, sacc);
                            }
                        } else if (has_stats) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (n0 + i < ep.N) { s1 += f[i]; s2 = fmaf(f[i], f[i], s2); }
                        }
                        if (tma_out) {
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmOut, st_main, n0, m_row0);
                                tma_store_commit();
                                if (has_copy) {
                                    tma_store_2d(&tmOut2, st_copy + (unit & 1) * (EPI_STAGE_BYTES / 2), n0, m_row0);
                                    tma_store_commit();
                                }
                            }
                        } else {
                            // rows only 8-byte aligned (LM head, ld = 32002): two rows x 128 B per warp instruction
                            __syncwarp();
                            const int rsub = lane >> 4, cp = (lane & 15) * 2;
                            float* og = reinterpret_cast<float*>(ep.out);
#pragma unroll 4
                            for (int rr = 0; rr < 32; rr += 2) {
                                const int r = rr + rsub;
                                const float2 val = *reinterpret_cast<const float2*>(st_main + r * 128 + (((cp >> 2) ^ (r & 7)) << 4) + (cp & 3) * 4);
                                const long long row = m_row0 + r;
                                if (row < ep.M) {
                                    float* dst = og + row * ep.ld_out + n0 + cp;
                                    if (n0 + cp + 1 < ep.N) *reinterpret_cast<float2*>(dst) = val;
                                    else if (n0 + cp < ep.N) *dst = val.x;
                                }
                            }
                            __syncwarp();
                        }
                        ++unit;
                    } else {
                        const int half = c & 1;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 pk;
                            pk.x = pack_bf16(f[8 * g + 0], f[8 * g + 1]);
                            pk.y = pack_bf16(f[8 * g + 2], f[8 * g + 3]);
                            pk.z = pack_bf16(f[8 * g + 4], f[8 * g + 5]);
                            pk.w = pack_bf16(f[8 * g + 6], f[8 * g + 7]);
                            *reinterpret_cast<uint4*>(mb + (((half * 4 + g) ^ sw) << 4)) = pk;
                            if (has_stats) stats_add8(pk, full, n0 + 8 * g, ep.N, sacc);
                        }
                        if (half == 1 || c == nchunks - 1) {
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmOut, st_main, nb + (c & ~1) * 32, m_row0);
                                tma_store_commit();
                            }
                            ++unit;
                        }
                    }
                }
                KX_GT(it, 0, 0 + 8 * 0);
                if (has_stats && m < ep.M && nb + c_begin * 32 < ep.N) {      // one partial per 128 columns
                    float t1, t2;
                    stats_fold(sacc, t1, t2);
                    ep.stats_out[static_cast<long long>(n_blk * 2 + chalf) * ep.M + m] = make_float2(s1 + t1, s2 + t2);
                }
                tc_fence_before();
                if constexpr (CG == 1) mbar_arrive(&tempty[a]); else mbar_arrive_cluster(&tempty[a], 0);
            }
            if (lane == 0) tma_store_wait<0>();
        }
    }

    __syncwarp();
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
}

// ----------------------------------------------------------------------------- host side
// cuTensorMapEncodeTiled costs 1-2 us of host time and a GEMM launch needs up to five maps; the un-graphed training step
// issues ~500 GEMMs over the same few hundred (buffer, shape) pairs every step, so encoded maps are cached per host
// thread, keyed by everything that goes into the encoding (the cache holds descriptors only — never device data).
struct TmapKey {
    const void* ptr; uint64_t inner, outer, stride; uint32_t box_inner, box_outer; int dtype, swizzle;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && stride == o.stride && box_inner == o.box_inner &&
               box_outer == o.box_outer && dtype == o.dtype && swizzle == o.swizzle;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        for (uint64_t v : {k.inner, k.outer, k.stride, (uint64_t(k.box_inner) << 32) | k.box_outer, (uint64_t(uint32_t(k.dtype)) << 32) | uint32_t(k.swizzle)})
            h = (h ^ v) * 0x100000001B3ull + (h >> 29);
        return static_cast<size_t>(h);
    }
};

static bool make_tmap_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer,
                         uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                         CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    const TmapKey key{ptr, inner, outer, row_stride_bytes, box_inner, box_outer, static_cast<int>(dtype), static_cast<int>(swizzle)};
    auto hit = cache.find(key);
    if (hit != cache.end()) { *tm = hit->second; return true; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    auto fn = driver_api().encode_tiled;
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver)"); return false; }
    CUresult r = fn(tm, dtype, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r, ptr,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
                  box_outer);
        return false;
    }
    if (cache.size() >= 8192) cache.clear();
    cache.emplace(key, *tm);
    return true;
}
bool make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer) {
    return make_tmap_2d(tm, ptr, inner, outer, row_stride_bytes, box_inner, box_outer);
}
bool make_tmap_f32_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
    return make_tmap_2d(tm, ptr, inner, outer, row_stride_bytes, box_inner, box_outer, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

template <int CG, int BN, bool OUT_F32, int EPI, bool TMA_EPI, bool ATR = false, bool BTR = false>
static int launch_gemm(const void* A, long long lda, const void* W, long long ldw, const GemmEpi& ep, int max_ctas,
                       cudaStream_t stream) {
    using Cfg = GemmCfg<CG, BN, TMA_EPI>;
    CUtensorMap tmA, tmB, tmOut, tmRes, tmOut2;
    if constexpr (!ATR) { if (!make_tmap_2d(&tmA, A, ep.K, ep.M, lda * 2, BLOCK_K, BLOCK_M)) return KX_ERR_TMAP; }
    else { if (!make_tmap_2d(&tmA, A, ep.M, ep.K, lda * 2, 64, BLOCK_K)) return KX_ERR_TMAP; }       // A given as [K, M]
    if constexpr (!BTR) { if (!make_tmap_2d(&tmB, W, ep.K, ep.N, ldw * 2, BLOCK_K, Cfg::B_ROWS)) return KX_ERR_TMAP; }
    else { if (!make_tmap_2d(&tmB, W, ep.N, ep.K, ldw * 2, 64, BLOCK_K)) return KX_ERR_TMAP; }       // W given as [K, N]
    if constexpr (TMA_EPI) {
        if (!ep.tma_store) {
            tmOut = tmA;                       // unused: the staged epilogue writes with coalesced float2 stores
        } else if (!make_tmap_2d(&tmOut, ep.out, ep.N, ep.M, ep.ld_out * (OUT_F32 ? 4 : 2), OUT_F32 ? 32 : 64, 32,
                                 OUT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16))
            return KX_ERR_TMAP;
        if (ep.res != nullptr) {
            if (!make_tmap_2d(&tmRes, ep.res, ep.N, ep.M, ep.ld_res * 4, 32, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32))
                return KX_ERR_TMAP;
        } else {
            tmRes = tmOut;
        }
        if (OUT_F32 && ep.out2 != nullptr) {
            if (!make_tmap_2d(&tmOut2, ep.out2, ep.N, ep.M, ep.ld_out2 * 2, 32, 32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              CU_TENSOR_MAP_SWIZZLE_64B))
                return KX_ERR_TMAP;
        } else {
            tmOut2 = tmOut;
        }
    } else {
        tmOut = tmA;
        tmRes = tmA;
        tmOut2 = tmA;
    }
    auto kern = gemm_bf16_kernel<CG, BN, OUT_F32, EPI, TMA_EPI, ATR, BTR>;
    static std::once_flag attr_once;   // per template instantiation; safe when several host threads launch GEMMs
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); });
    if (attr_err != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(attr_err)); return KX_ERR_LAUNCH; }
    const int num_m = (ep.M + BLOCK_M * CG - 1) / (BLOCK_M * CG);
    const int num_n = (ep.N + BN - 1) / BN;
    int clusters = std::min(num_m * num_n, std::max(1, max_ctas / CG));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CG);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmRes, tmOut2, ep);
    if (e != cudaSuccess) { set_error("gemm launch failed: %s", cudaGetErrorString(e)); return KX_ERR_LAUNCH; }
    count_launch();
    return KX_OK;
}

}  // namespace kx

using namespace kx;

#ifdef KX_GEMM_TRACE
extern "C" int kx_gemm_set_trace(long long* device_buffer) {
    cudaError_t e = cudaMemcpyToSymbol(g_gemm_trace, &device_buffer, sizeof(device_buffer));
    return e == cudaSuccess ? KX_OK : KX_ERR_LAUNCH;
}
#endif

extern "C" int kx_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const kx_gemm_args* g,
                            cudaStream_t stream) {
    if (!A || !W || !g || !g->out) { set_error("kx_gemm_bf16: null pointer"); return KX_ERR_ARG; }
    if (g->M <= 0 || g->N <= 0 || g->K <= 0) { set_error("kx_gemm_bf16: bad shape M=%d N=%d K=%d", g->M, g->N, g->K); return KX_ERR_ARG; }
    if ((lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) {
        set_error("kx_gemm_bf16: A/W need 16-byte aligned base and row pitch (lda=%lld ldw=%lld)", lda, ldw);
        return KX_ERR_ARG;
    }
    GemmEpi ep = {};
    ep.M = g->M; ep.N = g->N; ep.K = g->K;
    ep.raster_g = g->M >= g->N ? 8 : 16;      // tile order: see tile_coords
    ep.raster_nc = 8;
    ep.raster_nfast = 1;
    ep.bias = g->bias; ep.res = g->res; ep.ld_res = g->ld_res;
    ep.out = g->out; ep.ld_out = g->ld_out; ep.act = g->act;
    ep.grp_rows = g->grp_rows; ep.grp_stride = g->grp_stride; ep.grp_off = g->grp_off;
    ep.add_tab = g->add_tab; ep.add_off = g->add_off; ep.ld_add = g->ld_add;
    ep.xq_cos = g->xq_cos; ep.xq_sin = g->xq_sin; ep.xk_cos = g->xk_cos; ep.xk_sin = g->xk_sin;
    ep.seq_len = g->seq_len; ep.d_model = g->d_model;
    const int esz = g->out_f32 ? 4 : 2;
    bool vec = ((g->ld_out * esz) % 16 == 0) && ((reinterpret_cast<uintptr_t>(g->out) & 15) == 0);
    if (g->res) vec = vec && ((g->ld_res * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(g->res) & 15) == 0);
    if (g->add_tab) vec = vec && ((g->ld_add * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(g->add_tab) & 15) == 0);
    if (g->bias && (reinterpret_cast<uintptr_t>(g->bias) & 15)) { set_error("kx_gemm_bf16: bias must be 16-byte aligned"); return KX_ERR_ARG; }
    ep.vec_ok = vec ? 1 : 0;
    if (g->ln_part) {
        if (!g->ln_c || g->ln_tiles <= 0 || g->ln_cols <= 0 || (reinterpret_cast<uintptr_t>(g->ln_c) & 15) ||
            (reinterpret_cast<uintptr_t>(g->ln_part) & 7)) {
            set_error("kx_gemm_bf16: LayerNorm fold needs ln_c (16-byte aligned), ln_tiles > 0, ln_cols > 0");
            return KX_ERR_ARG;
        }
        ep.ln_part = reinterpret_cast<const float2*>(g->ln_part);
        ep.ln_c = g->ln_c;
        ep.ln_tiles = g->ln_tiles;
        ep.ln_inv_n = 1.0f / static_cast<float>(g->ln_cols);
        ep.ln_eps = g->ln_eps;
    }
    if (g->drop_p > 0.f) {
        if (!(g->drop_p < 1.f) || g->stats_out || g->ln_part || g->grp_rows) {
            set_error("kx_gemm_bf16: dropout needs 0 < p < 1 and the plain (unfolded, unscattered) epilogue of the training forward");
            return KX_ERR_ARG;
        }
        ep.drop = make_drop_spec(g->drop_p, g->drop_site, g->drop_seed);
    }
    ep.stats_out = reinterpret_cast<float2*>(g->stats_out);
    ep.out2 = g->out2;
    ep.ld_out2 = g->ld_out2;
    if (g->stats_out && (reinterpret_cast<uintptr_t>(g->stats_out) & 7)) { set_error("kx_gemm_bf16: stats_out must be 8-byte aligned"); return KX_ERR_ARG; }
    if (g->out2 && (!g->out_f32 || (g->ld_out2 % 8) || (reinterpret_cast<uintptr_t>(g->out2) & 15))) {
        set_error("kx_gemm_bf16: out2 (bf16 copy) needs an fp32 primary output and 16-byte aligned rows");
        return KX_ERR_ARG;
    }
    if (g->epi == KX_EPI_QKV_XPOS) {
        if (!g->xq_cos || !g->xq_sin || !g->xk_cos || !g->xk_sin || g->seq_len <= 0 || g->d_model <= 0 ||
            (g->d_model % 64) || g->N != 3 * g->d_model || g->out_f32) {
            set_error("kx_gemm_bf16: QKV_XPOS needs tables, seq_len, d_model%%64==0, N==3*d_model, bf16 out");
            return KX_ERR_ARG;
        }
    }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    int cg = g->cta_group;
    if (cg == 0) cg = (g->M > 128) ? 2 : 1;
    int bn = g->block_n;
    if (g->stats_out && bn == 0) bn = 256;           // partials are per half tile: 128 columns at BN=256, 64 at BN=128
    if (bn == 0) bn = (g->N >= 256 && (long long)((g->M + 127) / 128) * ((g->N + 255) / 256) >= sms / 2) ? 256 : 128;
    const int max_ctas = g->max_ctas > 0 ? std::min(g->max_ctas, sms) : sms;

    // staged TMA epilogue whenever the output (and residual) rows are 16-byte aligned and unscattered;
    // otherwise (LM head with ld = 32002, image_proj / patch-embed row scatter) direct predicated stores
    bool tma_epi = (g->grp_rows == 0) && (g->add_tab == nullptr) && ((g->ld_out * esz) % 16 == 0) &&
                   ((reinterpret_cast<uintptr_t>(g->out) & 15) == 0);
    if (g->res) tma_epi = tma_epi && ((g->ld_res * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(g->res) & 15) == 0);
    if (g->res && !g->out_f32) tma_epi = false;       // (the staged bf16 store path carries no residual)
    if (g->epi_mode == 1) tma_epi = false;
    ep.tma_store = tma_epi ? 1 : 0;
    // fp32 rows that are only 8-byte aligned (LM head: ld = 32002) still take the staged kernel, which then writes
    // two 128-byte row segments per warp instruction instead of 32 scattered 8-byte pieces
    if (!tma_epi && g->epi_mode != 1 && g->out_f32 && !g->res && !g->out2 && !g->stats_out && g->grp_rows == 0 &&
        g->add_tab == nullptr && (g->ld_out % 2 == 0) && ((reinterpret_cast<uintptr_t>(g->out) & 7) == 0))
        tma_epi = true;
    if ((g->stats_out || g->out2) && !ep.tma_store) {
        set_error("kx_gemm_bf16: stats_out / out2 need the staged epilogue (unscattered, 16-byte aligned rows, epi_mode 0)");
        return KX_ERR_ARG;
    }
    if (g->a_trans || g->b_trans) {
        // backward GEMMs: dgrad (W transposed) and wgrad (both transposed); staged epilogue only
        if (g->a_trans && !g->b_trans) { set_error("kx_gemm_bf16: a_trans without b_trans is not built"); return KX_ERR_ARG; }
        if (g->epi != KX_EPI_GENERIC || !tma_epi || !ep.tma_store) {
            set_error("kx_gemm_bf16: transposed operands need the generic staged epilogue (16-byte aligned, unscattered output)");
            return KX_ERR_ARG;
        }
        const bool big = (cg == 2 && bn == 256);
        if (!big) { cg = 1; bn = 128; }
#define KX_GEMM_TR(CG_, BN_, ATR_)                                                                                                  \
        {                                                                                                                           \
            if (g->out_f32) return launch_gemm<CG_, BN_, true, KX_EPI_GENERIC, true, ATR_, true>(A, lda, W, ldw, ep, max_ctas, stream); \
            return launch_gemm<CG_, BN_, false, KX_EPI_GENERIC, true, ATR_, true>(A, lda, W, ldw, ep, max_ctas, stream);               \
        }
        if (big) { if (g->a_trans) KX_GEMM_TR(2, 256, true) else KX_GEMM_TR(2, 256, false) }
        else { if (g->a_trans) KX_GEMM_TR(1, 128, true) else KX_GEMM_TR(1, 128, false) }
#undef KX_GEMM_TR
    }
#define KX_GEMM_CASE2(CG_, BN_, T_)                                                                          \
    {                                                                                                        \
        if (g->epi == KX_EPI_QKV_XPOS) return launch_gemm<CG_, BN_, false, KX_EPI_QKV_XPOS, T_>(A, lda, W, ldw, ep, max_ctas, stream); \
        if (g->out_f32) return launch_gemm<CG_, BN_, true, KX_EPI_GENERIC, T_>(A, lda, W, ldw, ep, max_ctas, stream);                  \
        return launch_gemm<CG_, BN_, false, KX_EPI_GENERIC, T_>(A, lda, W, ldw, ep, max_ctas, stream);                                  \
    }
#define KX_GEMM_CASE(CG_, BN_)                                                                               \
    if (cg == CG_ && bn == BN_) {                                                                            \
        if (tma_epi) KX_GEMM_CASE2(CG_, BN_, true) else KX_GEMM_CASE2(CG_, BN_, false)                       \
    }
    KX_GEMM_CASE(1, 128)
    KX_GEMM_CASE(1, 256)
    KX_GEMM_CASE(2, 128)
    KX_GEMM_CASE(2, 256)
#undef KX_GEMM_CASE
#undef KX_GEMM_CASE2
    set_error("kx_gemm_bf16: unsupported config cta_group=%d block_n=%d", cg, bn);
    return KX_ERR_ARG;
}
