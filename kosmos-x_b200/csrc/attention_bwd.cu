// kx_attn_bwd: flash-attention backward, head_dim 64, tcgen05 / TMEM.
//
// Replaces autograd through torchscale MultiheadAttention's bmm -> +triu(-inf) -> softmax(fp32) -> bmm core
//  in the training step (§8(a) a19).  Nothing of size T x T touches HBM.
//
// One CTA per (batch, head, 128-key block j); it walks the query blocks i that attend to those keys
// (i >= j when causal).  Everything is computed TRANSPOSED (rows = keys) so that the probability and
// score-gradient tiles land in TMEM with one key per lane and can feed the next MMAs as TMEM A operands:
//   S^T  = K_j . Q_i^T            SS   (A = K tile, B = Q tile, both K-major)                 -> TMEM [0,128)
//   dP^T = V_j . dO_i^T           SS                                                          -> TMEM [128,256)
//   P^T  = exp2(S^T*scale*log2e - lse_i)   dS^T = P^T * (dP^T - delta_i)                      (registers)
//   dV_j += P^T . dO_i            TS   (A = P^T bf16 in TMEM over S^T, B = dO tile MN-major)  -> TMEM [256,320)
//   dK_j += dS^T . Q_i            TS   (A = dS^T bf16 in TMEM over dP^T, B = Q tile MN-major) -> TMEM [320,384)
//   dQ_i  = dS . K_j              SS   (A = dS^T staged in smem, read MN-major; B = K tile MN-major) -> TMEM [384,448)
// dQ_i is added to an fp32 accumulator in global memory with vector atomics (the only cross-CTA reduction);
// dK_j / dV_j leave through the epilogue as bf16.  `scale` is applied on the way out.
//   warps 0-7  compute: thread = key row (TMEM lane), warpgroup g owns query columns [64g, 64g+64)
//   warp 8     TMA producer: K_j, V_j once; Q_i, dO_i, lse_i, delta_i through a 2-slot ring
//   warp 9     MMA issuer (one elected thread)
#include "kx_internal.h"
#include "ptx.cuh"

namespace kx {

constexpr int BW_THREADS = 384;
constexpr int BW_TILE = 128 * 64 * 2;                 // [128 x 64] bf16
constexpr int BW_SMEM_K = 0;
constexpr int BW_SMEM_V = BW_SMEM_K + BW_TILE;
constexpr int BW_SMEM_Q = BW_SMEM_V + BW_TILE;        // 2 slots
constexpr int BW_SMEM_DO = BW_SMEM_Q + 2 * BW_TILE;   // 2 slots
constexpr int BW_SMEM_DS = BW_SMEM_DO + 2 * BW_TILE;  // dS^T: 2 MN atoms of [128 keys][64 queries]
constexpr int BW_SMEM_LSE = BW_SMEM_DS + 2 * BW_TILE; // [2][128] fp32
constexpr int BW_SMEM_DELTA = BW_SMEM_LSE + 2 * 512;
constexpr int BW_SMEM_BAR = BW_SMEM_DELTA + 2 * 512;
constexpr int BW_SMEM_BYTES = BW_SMEM_BAR + 128;
constexpr int BW_TMEM_COLS = 512;
constexpr int BW_T_S = 0, BW_T_DP = 128, BW_T_DV = 256, BW_T_DK = 320, BW_T_DQ = 384;

struct AttnBwdParams {
    const float* lse;          // [heads][batch][t_pad], log2 units (kx_attn_fwd_lse)
    const float* delta;        // [heads][batch][t_pad], rowsum(dO * O)
    float* dq_accum;           // fp32 [batch*seq_len, heads*64]
    __nv_bfloat16* dk;
    __nv_bfloat16* dv;
    long long ld_dkv;
    int seq_len, heads, batch, t_pad;
    float scale, scale_log2;
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <bool CAUSAL>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_SMEM_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* full = bars + 1;        // [2] Q/dO/lse/delta slot filled
    uint64_t* empty = bars + 3;       // [2] slot consumed by the MMAs
    uint64_t* bar_s = bars + 5;       // S^T complete
    uint64_t* bar_dp = bars + 6;      // dP^T complete
    uint64_t* bar_pds = bars + 7;     // P^T, dS^T written (256 compute threads)
    uint64_t* bar_dq = bars + 8;      // dV, dK, dQ MMAs of the iteration complete
    uint64_t* bar_dqr = bars + 9;     // dQ read out of TMEM (256 compute threads)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    float* s_lse = reinterpret_cast<float*>(smem + BW_SMEM_LSE);
    float* s_delta = reinterpret_cast<float*>(smem + BW_SMEM_DELTA);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = p.seq_len;
    const int nblk = (T + 127) >> 7;
    const int bh_count = p.heads * p.batch;
    const int j = blockIdx.x / bh_count;                 // key block; heavy (small j) first when causal
    const int bh = blockIdx.x - j * bh_count;
    const int head = bh % p.heads, b = bh / p.heads;
    const int row_base = b * T;
    const int i0 = CAUSAL ? j : 0;
    const int n_it = nblk - i0;
    const long long vec_base = (static_cast<long long>(head) * p.batch + b) * p.t_pad;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023) { printf("kx attn_bwd: dynamic smem base not 1024-aligned\n"); __trap(); }
        mbar_init(kv_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_pds, 256); mbar_init(bar_dq, 1); mbar_init(bar_dqr, 256);
        fence_mbar_init();
    }
    if (warp == 8) {
        if (lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmDO); }
        tmem_alloc<1>(tmem_slot, BW_TMEM_COLS);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 8) {
            if (elect_one()) {
                // ================= TMA producer =================
                mbar_arrive_expect_tx(kv_full, 2 * BW_TILE);
                tma_load_2d(&tmK, kv_full, smem + BW_SMEM_K, head * 64, row_base + j * 128, kEvictFirst);
                tma_load_2d(&tmV, kv_full, smem + BW_SMEM_V, head * 64, row_base + j * 128, kEvictFirst);
                for (int it = 0; it < n_it; ++it) {
                    const int i = i0 + it, s = it & 1;
                    mbar_wait(&empty[s], ((it >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full[s], 2 * BW_TILE + 2 * 512);
                    tma_load_2d(&tmQ, &full[s], smem + BW_SMEM_Q + s * BW_TILE, head * 64, row_base + i * 128, kEvictLast);
                    tma_load_2d(&tmDO, &full[s], smem + BW_SMEM_DO + s * BW_TILE, head * 64, row_base + i * 128, kEvictLast);
                    bulk_load_1d(s_lse + s * 128, p.lse + vec_base + i * 128, 512, &full[s]);
                    bulk_load_1d(s_delta + s * 128, p.delta + vec_base + i * 128, 512, &full[s]);
                }
            }
        } else if (warp == 9) {
            if (elect_one()) {
                // ================= MMA issuer =================
                constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);    // [keys x queries] = A(K-major) . B(K-major)^T
                constexpr uint32_t idesc_kv = make_idesc_bf16(128, 64, 0, 1);    // [keys x d] = A(TMEM) . B(MN-major)
                constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);    // [queries x d] = A(MN-major smem) . B(MN-major)
                const uint32_t t_s = tmem_base + BW_T_S, t_dp = tmem_base + BW_T_DP, t_dv = tmem_base + BW_T_DV,
                               t_dk = tmem_base + BW_T_DK, t_dq = tmem_base + BW_T_DQ;
                const uint64_t k_kmaj = make_desc_sw128(smem_u32(smem + BW_SMEM_K));
                const uint64_t v_kmaj = make_desc_sw128(smem_u32(smem + BW_SMEM_V));
                const uint64_t k_mn = make_desc_sw128(smem_u32(smem + BW_SMEM_K), BW_TILE);
                const uint64_t ds_mn = make_desc_sw128(smem_u32(smem + BW_SMEM_DS), BW_TILE);   // LBO = 16 KB between the two query atoms
                auto issue_scores = [&](int s) {
                    const uint64_t q_kmaj = make_desc_sw128(smem_u32(smem + BW_SMEM_Q + s * BW_TILE));
                    const uint64_t do_kmaj = make_desc_sw128(smem_u32(smem + BW_SMEM_DO + s * BW_TILE));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16<1>(t_s, k_kmaj + 2 * k, q_kmaj + 2 * k, idesc_s, k != 0);
                    umma_commit(bar_s);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16<1>(t_dp, v_kmaj + 2 * k, do_kmaj + 2 * k, idesc_s, k != 0);
                    umma_commit(bar_dp);
                };
                mbar_wait(kv_full, 0);
                mbar_wait(&full[0], 0);
                tc_fence_after();
                issue_scores(0);
                for (int it = 0; it < n_it; ++it) {
                    const int s = it & 1;
                    const uint64_t q_mn = make_desc_sw128(smem_u32(smem + BW_SMEM_Q + s * BW_TILE), BW_TILE);
                    const uint64_t do_mn = make_desc_sw128(smem_u32(smem + BW_SMEM_DO + s * BW_TILE), BW_TILE);
                    mbar_wait(bar_pds, it & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k)          // 16 queries per step: 8 TMEM columns of bf16 pairs, 16 dO rows = 2 KB
                        umma_bf16_ts(t_dv, t_s + k * 8, do_mn + k * 128, idesc_kv, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_bf16_ts(t_dk, t_dp + k * 8, q_mn + k * 128, idesc_kv, (it > 0 || k > 0) ? 1u : 0u);
                    if (it > 0) {                        // dQ of the previous iteration has been read out of TMEM
                        mbar_wait(bar_dqr, (it - 1) & 1);
                        tc_fence_after();
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)          // 16 keys per step: 16 rows of dS^T / K = 2 KB
                        umma_bf16<1>(t_dq, ds_mn + k * 128, k_mn + k * 128, idesc_dq, k != 0);
                    umma_commit(bar_dq);
                    umma_commit(&empty[s]);
                    if (it + 1 < n_it) {
                        mbar_wait(&full[s ^ 1], ((it + 1) >> 1) & 1);
                        tc_fence_after();
                        issue_scores(s ^ 1);
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ================= compute: thread = key row r (TMEM lane), warpgroup g = query column half =================
        const int g = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int kg = j * 128 + r;
        const float sl2 = p.scale_log2;
        uint8_t* ds_row = smem + BW_SMEM_DS + g * BW_TILE + r * 128;
        const int sw = r & 7;

        for (int it = 0; it < n_it; ++it) {
            const int i = i0 + it, s = it & 1;
            mbar_wait(&full[s], (it >> 1) & 1);                  // lse / delta of this query block are in smem
            mbar_wait(bar_s, it & 1);
            tc_fence_after();
            uint32_t sv[64], dpv[64];
            tmem_ld32(tmem_base + lane_addr + BW_T_S + g * 64, reinterpret_cast<uint32_t(&)[32]>(sv[0]));
            tmem_ld32(tmem_base + lane_addr + BW_T_S + g * 64 + 32, reinterpret_cast<uint32_t(&)[32]>(sv[32]));
            mbar_wait(bar_dp, it & 1);
            tc_fence_after();
            tmem_ld32(tmem_base + lane_addr + BW_T_DP + g * 64, reinterpret_cast<uint32_t(&)[32]>(dpv[0]));
            tmem_ld32(tmem_base + lane_addr + BW_T_DP + g * 64 + 32, reinterpret_cast<uint32_t(&)[32]>(dpv[32]));
            tmem_ld_wait();
            tc_fence_before();
            named_bar_sync(1, 256);                              // both warpgroups hold S^T / dP^T: the regions may be overwritten

            const float* lse = s_lse + s * 128 + g * 64;
            const float* dl = s_delta + s * 128 + g * 64;
            const bool edge = (CAUSAL && i == j) || (i * 128 + 128 > T) || (j * 128 + 128 > T);
            uint32_t pp[32], dd[32];
#pragma unroll
            for (int c = 0; c < 64; c += 2) {
                const float2 l2 = *reinterpret_cast<const float2*>(lse + c);
                const float2 d2 = *reinterpret_cast<const float2*>(dl + c);
                float p0 = ex2_approx(fmaf(__uint_as_float(sv[c]), sl2, -l2.x));
                float p1 = ex2_approx(fmaf(__uint_as_float(sv[c + 1]), sl2, -l2.y));
                if (edge) {
                    const int q0 = i * 128 + g * 64 + c;
                    const bool ok0 = kg < T && q0 < T && (!CAUSAL || q0 >= kg);
                    const bool ok1 = kg < T && q0 + 1 < T && (!CAUSAL || q0 + 1 >= kg);
                    p0 = ok0 ? p0 : 0.f;
                    p1 = ok1 ? p1 : 0.f;
                }
                float ds0 = p0 * (__uint_as_float(dpv[c]) - d2.x);
                float ds1 = p1 * (__uint_as_float(dpv[c + 1]) - d2.y);
                if (edge) {                                       // garbage rows beyond T may hold inf/nan in dP or delta
                    ds0 = (p0 == 0.f) ? 0.f : ds0;
                    ds1 = (p1 == 0.f) ? 0.f : ds1;
                }
                pp[c >> 1] = pack_bf16(p0, p1);
                dd[c >> 1] = pack_bf16(ds0, ds1);
            }
            tmem_st32(tmem_base + lane_addr + BW_T_S + g * 32, pp);        // P^T over S^T (bf16 pairs)
            tmem_st32(tmem_base + lane_addr + BW_T_DP + g * 32, dd);       // dS^T over dP^T
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)                                  // dS^T row -> smem atom g, SWIZZLE_128B
                *reinterpret_cast<uint4*>(ds_row + ((ch ^ sw) << 4)) = make_uint4(dd[4 * ch], dd[4 * ch + 1], dd[4 * ch + 2], dd[4 * ch + 3]);
            tmem_st_wait();
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar_pds);

            // ---- dQ_i (rows = queries now): TMEM -> fp32 accumulator in global memory
            mbar_wait(bar_dq, it & 1);
            tc_fence_after();
            uint32_t qv[32];
            tmem_ld32(tmem_base + lane_addr + BW_T_DQ + g * 32, qv);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_dqr);
            const int qg = i * 128 + r;
            if (qg < T) {
                float4* dst = reinterpret_cast<float4*>(p.dq_accum + static_cast<long long>(row_base + qg) * (p.heads * 64) + head * 64 + g * 32);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    atomicAdd(dst + u, make_float4(__uint_as_float(qv[4 * u]) * p.scale, __uint_as_float(qv[4 * u + 1]) * p.scale,
                                                   __uint_as_float(qv[4 * u + 2]) * p.scale, __uint_as_float(qv[4 * u + 3]) * p.scale));
            }
        }
        // ---- epilogue: dV_j, dK_j (bar_dq of the last iteration covers every MMA)
        tc_fence_after();
        uint32_t vv[32], kk[32];
        tmem_ld32(tmem_base + lane_addr + BW_T_DV + g * 32, vv);
        tmem_ld32(tmem_base + lane_addr + BW_T_DK + g * 32, kk);
        tmem_ld_wait();
        if (kg < T) {
            const long long off = static_cast<long long>(row_base + kg) * p.ld_dkv + head * 64 + g * 32;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint4 a, c;
                a.x = pack_bf16(__uint_as_float(vv[8 * u]), __uint_as_float(vv[8 * u + 1]));
                a.y = pack_bf16(__uint_as_float(vv[8 * u + 2]), __uint_as_float(vv[8 * u + 3]));
                a.z = pack_bf16(__uint_as_float(vv[8 * u + 4]), __uint_as_float(vv[8 * u + 5]));
                a.w = pack_bf16(__uint_as_float(vv[8 * u + 6]), __uint_as_float(vv[8 * u + 7]));
                c.x = pack_bf16(__uint_as_float(kk[8 * u]) * p.scale, __uint_as_float(kk[8 * u + 1]) * p.scale);
                c.y = pack_bf16(__uint_as_float(kk[8 * u + 2]) * p.scale, __uint_as_float(kk[8 * u + 3]) * p.scale);
                c.z = pack_bf16(__uint_as_float(kk[8 * u + 4]) * p.scale, __uint_as_float(kk[8 * u + 5]) * p.scale);
                c.w = pack_bf16(__uint_as_float(kk[8 * u + 6]) * p.scale, __uint_as_float(kk[8 * u + 7]) * p.scale);
                *reinterpret_cast<uint4*>(p.dv + off + 8 * u) = a;
                *reinterpret_cast<uint4*>(p.dk + off + 8 * u) = c;
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<1>(tmem_base, BW_TMEM_COLS);
}

// delta[h][b][t] = sum_d dO[b,t,h,d] * O[b,t,h,d]   (8 lanes per (row, head), 16-byte loads)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ld_o, const __nv_bfloat16* __restrict__ d_o, long long ld_do,
                  float* __restrict__ delta, int batch, int heads, int seq_len, int t_pad) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long total = static_cast<long long>(batch) * seq_len * heads * 8;
    const bool active = idx < total;
    const long long ci = active ? idx : 0;
    const int sub = static_cast<int>(ci & 7);
    const long long rh = ci >> 3;
    const int head = static_cast<int>(rh % heads);
    const long long row = rh / heads;
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * ld_o + head * 64 + sub * 8);
    const uint4 c = *reinterpret_cast<const uint4*>(d_o + row * ld_do + head * 64 + sub * 8);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float2 x = __bfloat1622float2(ha[u]), y = __bfloat1622float2(hc[u]);
        s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (active && sub == 0) {
        const int b = static_cast<int>(row / seq_len), t = static_cast<int>(row - static_cast<long long>(b) * seq_len);
        delta[(static_cast<long long>(head) * batch + b) * t_pad + t] = s;
    }
}

// dq (bf16) = R^T(dq_accum) and dk = R^T(dk) in place, R = the xPos rotation of the QKV epilogue (kx_xpos_bwd does the
// same on bf16 inputs; this variant reads the fp32 dQ accumulator).  ld in elements of the bf16 matrix.
__global__ void __launch_bounds__(256)
attn_bwd_finish_kernel(const float* __restrict__ dq_accum, __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk,
                       long long ld, int rows, int d_model, int seq_len, const float* __restrict__ q_cos,
                       const float* __restrict__ q_sin, const float* __restrict__ k_cos, const float* __restrict__ k_sin) {
    const int vec_per_row = (2 * d_model) >> 3;
    const long long total = static_cast<long long>(rows) * vec_per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / vec_per_row);
        int col = static_cast<int>(i - static_cast<long long>(row) * vec_per_row) * 8;
        const bool is_k = col >= d_model;
        if (is_k) col -= d_model;
        float v[8];
        if (!is_k) {
            const float4 a = *reinterpret_cast<const float4*>(dq_accum + static_cast<long long>(row) * d_model + col);
            const float4 c = *reinterpret_cast<const float4*>(dq_accum + static_cast<long long>(row) * d_model + col + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        } else {
            const uint4 q = *reinterpret_cast<const uint4*>(dk + static_cast<long long>(row) * ld + col);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int u = 0; u < 4; ++u) { const float2 f = __bfloat1622float2(h[u]); v[2 * u] = f.x; v[2 * u + 1] = f.y; }
        }
        float o[8];
        if (q_cos != nullptr) {
            const int t = row % seq_len;
            const int j0 = (col & 63) >> 1;
            const float4 c = __ldg(reinterpret_cast<const float4*>((is_k ? k_cos : q_cos) + t * 32 + j0));
            const float4 s = __ldg(reinterpret_cast<const float4*>((is_k ? k_sin : q_sin) + t * 32 + j0));
            const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                o[2 * u] = v[2 * u] * cc[u] + v[2 * u + 1] * ss[u];
                o[2 * u + 1] = v[2 * u + 1] * cc[u] - v[2 * u] * ss[u];
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) o[u] = v[u];
        }
        uint4 q;
        q.x = pack_bf16(o[0], o[1]); q.y = pack_bf16(o[2], o[3]); q.z = pack_bf16(o[4], o[5]); q.w = pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>((is_k ? dk : dq) + static_cast<long long>(row) * ld + col) = q;
    }
}

}  // namespace kx

using namespace kx;

extern "C" int kx_attn_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                           const void* d_out, long long ld_dout, const float* lse, void* dq, void* dk, void* dv,
                           long long ld_dqkv, float* dq_accum, float* delta, const float* xq_cos, const float* xq_sin,
                           const float* xk_cos, const float* xk_sin, int batch, int heads, int seq_len, int causal, float scale,
                           cudaStream_t stream) {
    if (!q || !k || !v || !out || !d_out || !lse || !dq || !dk || !dv || !dq_accum || !delta) { set_error("kx_attn_bwd: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || heads <= 0 || seq_len <= 0 || (ld_qkv % 8) || (ld_out % 8) || (ld_dout % 8) || (ld_dqkv % 8)) {
        set_error("kx_attn_bwd: bad shape / pitch (16-byte aligned rows required)");
        return KX_ERR_ARG;
    }
    for (const void* ptr : {q, k, v, out, d_out, (const void*)dq, (const void*)dk, (const void*)dv, (const void*)dq_accum,
                            (const void*)delta, (const void*)lse})
        if (reinterpret_cast<uintptr_t>(ptr) & 15) { set_error("kx_attn_bwd: pointers must be 16-byte aligned"); return KX_ERR_ARG; }
    const bool rot = xq_cos || xq_sin || xk_cos || xk_sin;
    if (rot && !(xq_cos && xq_sin && xk_cos && xk_sin)) { set_error("kx_attn_bwd: give all four xPos tables or none"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const unsigned long long rows = static_cast<unsigned long long>(batch) * seq_len;
    const int nblk = (seq_len + 127) / 128;
    const int t_pad = nblk * 128;
    CUtensorMap tq, tk, tv, tdo;
    if (!make_tmap_bf16_2d(&tq, q, (uint64_t)heads * 64, rows, ld_qkv * 2, 64, 128)) return KX_ERR_TMAP;
    if (!make_tmap_bf16_2d(&tk, k, (uint64_t)heads * 64, rows, ld_qkv * 2, 64, 128)) return KX_ERR_TMAP;
    if (!make_tmap_bf16_2d(&tv, v, (uint64_t)heads * 64, rows, ld_qkv * 2, 64, 128)) return KX_ERR_TMAP;
    if (!make_tmap_bf16_2d(&tdo, d_out, (uint64_t)heads * 64, rows, ld_dout * 2, 64, 128)) return KX_ERR_TMAP;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES);
        cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("kx_attn_bwd: cudaFuncSetAttribute failed"); return KX_ERR_LAUNCH; }
        attr_set = true;
    }
    const long long d_model = static_cast<long long>(heads) * 64;
    cudaError_t e = cudaMemsetAsync(dq_accum, 0, rows * d_model * sizeof(float), stream);
    if (e != cudaSuccess) { set_error("kx_attn_bwd: memset failed: %s", cudaGetErrorString(e)); return KX_ERR_LAUNCH; }
    {
        const long long total = static_cast<long long>(rows) * heads * 8;
        attn_delta_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(out), ld_out, reinterpret_cast<const __nv_bfloat16*>(d_out), ld_dout, delta,
            batch, heads, seq_len, t_pad);
        int st = check_launch("kx_attn_bwd (delta)");
        if (st != KX_OK) return st;
    }
    AttnBwdParams p;
    p.lse = lse; p.delta = delta; p.dq_accum = dq_accum;
    p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.ld_dkv = ld_dqkv;
    p.seq_len = seq_len; p.heads = heads; p.batch = batch; p.t_pad = t_pad;
    p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
    const long long ctas = static_cast<long long>(nblk) * heads * batch;
    if (ctas > 0x7fffffffLL) { set_error("kx_attn_bwd: too many tiles"); return KX_ERR_ARG; }
    if (causal) attn_bwd_kernel<true><<<static_cast<unsigned>(ctas), BW_THREADS, BW_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, p);
    else attn_bwd_kernel<false><<<static_cast<unsigned>(ctas), BW_THREADS, BW_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, p);
    int st = check_launch("kx_attn_bwd");
    if (st != KX_OK) return st;
    {
        const long long total = static_cast<long long>(rows) * (2 * d_model / 8);
        const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sms) * 8));
        attn_bwd_finish_kernel<<<blocks, 256, 0, stream>>>(dq_accum, reinterpret_cast<__nv_bfloat16*>(dq),
                                                          reinterpret_cast<__nv_bfloat16*>(dk), ld_dqkv, static_cast<int>(rows),
                                                          static_cast<int>(d_model), seq_len, xq_cos, xq_sin, xk_cos, xk_sin);
        st = check_launch("kx_attn_bwd (finish)");
    }
    return st;
}
