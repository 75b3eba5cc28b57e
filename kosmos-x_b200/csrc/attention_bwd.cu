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
//   dV_j += P^T . dO_i            TS   (A = P^T bf16 in TMEM [448,512), B = dO tile MN-major)  -> TMEM [256,320)
//   dK_j += dS^T . Q_i            SS   (A = dS^T staged in smem, read K-major;  B = Q tile MN-major) -> TMEM [320,384)
//   dQ_i  = dS . K_j              SS   (A = the same smem bytes, read MN-major; B = K tile MN-major) -> TMEM [384,448)
// Neither S^T nor dP^T is overwritten by the compute warps, so the tensor core computes the scores of block i+1 while
// block i is still in the softmax math (S^T / dP^T are released as soon as they are in registers).
// dQ_i is added to an fp32 accumulator in global memory by TMA reduction (cp.reduce.async.bulk.tensor .add from a
// swizzled smem staging tile: the only cross-CTA reduction, off the SM's load/store path);
// dK_j / dV_j leave through the epilogue as bf16.  `scale` is applied on the way out.
//   warps 0-15 compute: thread = key row (TMEM lane), warpgroup g owns query columns [32g, 32g+32); the per-element
//              work is packed f32x2: one LDS.128 brings (-lse, -delta) of a query pair, then FFMA2, 2 x MUFU.EX2, FADD2,
//              FMUL2 and two bf16x2 packs (the first version spent 13 instructions per element and was issue-bound
//              with 2 warps per scheduler: ncu showed the tensor pipe 19 % busy)
//   warp 16    TMA producer: K_j, V_j once; Q_i, dO_i and the (-lse, -delta) rows of block i through a 2-slot ring
//   warp 17    MMA issuer (one elected thread).  Issue order per iteration: dV_i, S^T_{i+1}, dK_i, dP^T_{i+1}, dQ_i — the
//              next block's scores are ready while the compute warps are still storing, and dQ leaves the critical path
//   warps 20-23 dQ drain: TMEM -> swizzled smem staging -> TMA reduce-add, concurrent with the next block's softmax math
//              (clock64 timeline, profiles/r1_attn_bwd_trace_*.txt: with the compute warps draining dQ themselves an
//              iteration cost 4840 cycles, 1450 of them in the drain)
#include "kx_internal.h"
#include "ptx.cuh"

namespace kx {

constexpr int BW_THREADS = 768;                       // 6 warpgroups: 4 compute, 1 {TMA, MMA, 2 idle warps}, 1 dQ drain
constexpr int BW_COMPUTE = 512;
constexpr int BW_W_TMA = 16, BW_W_MMA = 17, BW_W_DRAIN = 20;
constexpr int BW_TILE = 128 * 64 * 2;                 // [128 x 64] bf16
constexpr int BW_SMEM_K = 0;
constexpr int BW_SMEM_V = BW_SMEM_K + BW_TILE;
constexpr int BW_SLOTS = 3;                           // Q / dO / (-lse, -delta) ring: a slot is released at the END of its iteration
                                                      // (dK, dV read it last), so two slots exposed the whole TMA latency every block
constexpr int BW_SMEM_Q = BW_SMEM_V + BW_TILE;
constexpr int BW_SMEM_DO = BW_SMEM_Q + BW_SLOTS * BW_TILE;
constexpr int BW_SMEM_DS = BW_SMEM_DO + BW_SLOTS * BW_TILE;  // dS^T: 2 MN atoms of [128 keys][64 queries]
constexpr int BW_SMEM_DQ = BW_SMEM_DS + 2 * BW_TILE;  // dQ staging: 2 column halves of [128 queries][32 fp32], SWIZZLE_128B
constexpr int BW_SMEM_LD = BW_SMEM_DQ + 2 * BW_TILE;  // [slots][128 queries] (-lse, -delta) fp32 pairs
constexpr int BW_SMEM_BAR = BW_SMEM_LD + BW_SLOTS * 1024;
static_assert(BW_SMEM_BAR + 128 <= 227 * 1024, "shared memory budget");
constexpr int BW_SMEM_BYTES = BW_SMEM_BAR + 128;
constexpr int BW_TMEM_COLS = 512;
constexpr int BW_T_S = 0, BW_T_DP = 128, BW_T_DV = 256, BW_T_DK = 320, BW_T_DQ = 384, BW_T_P = 448;

struct AttnBwdParams {
    const float2* nld;         // [heads][batch][t_pad] pairs (-lse in log2 units, -rowsum(dO * O)), from attn_delta_kernel
    float* dq_accum;           // fp32 [batch*seq_len, heads*64]
    __nv_bfloat16* dk;
    __nv_bfloat16* dv;
    long long ld_dkv;
    int seq_len, heads, batch, t_pad;
    float scale, scale_log2;
    long long* trace;          // TRACE builds only: clock64 stamps of one CTA, [role: 0 compute thread 0, 1 MMA thread][iteration][16 points], then per-CTA records
    const uint32_t* drop_mask; // DROP builds only: key-major keep bits of kx_attn_dropout_masks, word per (tile, query quarter, key)
    float inv_keep;            // 1 / (1 - p)
    const float* k_cos;        // [seq_len][32] xPos tables of the keys (transpose of the QKV epilogue's rotation, applied to dK on
    const float* k_sin;        // the way out), or null
};

static long long* g_attn_bwd_trace = nullptr;         // kx_attn_bwd_set_trace

#define KX_BT(role, iter, point)                                                                      \
    do {                                                                                              \
        if constexpr (TRACE) {                                                                        \
            if (p.trace != nullptr && blockIdx.x == trace_cta && (iter) < 32)                         \
                p.trace[((role) * 32 + (iter)) * 16 + (point)] = clock64();                           \
        }                                                                                             \
    } while (0)

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <bool CAUSAL, bool TRACE = false, bool DROP = false>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_SMEM_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* full = bars + 1;        // [BW_SLOTS] Q/dO/lse/delta slot filled
    uint64_t* empty = bars + 4;       // [BW_SLOTS] slot consumed by the MMAs
    uint64_t* bar_s = bars + 7;       // S^T complete
    uint64_t* bar_dp = bars + 8;      // dP^T complete
    uint64_t* bar_pds = bars + 9;     // P^T, dS^T written (compute threads)
    uint64_t* bar_dq = bars + 10;     // dV, dK, dQ MMAs of the iteration complete
    uint64_t* bar_dqr = bars + 11;    // dQ read out of TMEM (128 drain threads)
    uint64_t* bar_fin = bars + 12;    // every MMA of the CTA complete (epilogue may read dV, dK)
    uint64_t* bar_sfree = bars + 13;  // S^T and dP^T are in registers (compute threads): the next scores may be issued
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    float4* s_ld = reinterpret_cast<float4*>(smem + BW_SMEM_LD);          // [2][64] {-lse0, -delta0, -lse1, -delta1}

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = p.seq_len;
    const int nblk = (T + 127) >> 7;
    // TRACE builds: the CTA whose timeline is recorded is named by the buffer's last timeline word (a negative value means
    // CTA 0); every CTA also leaves {SM id, globaltimer at entry, at its first ready score tile, at exit} behind the timelines.
    [[maybe_unused]] unsigned trace_cta = 0;
    if constexpr (TRACE) {
        if (p.trace != nullptr) {
            const long long want = p.trace[1023];
            trace_cta = want > 0 ? static_cast<unsigned>(want) : 0u;
            if (threadIdx.x == 0) {
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                p.trace[1024 + 4 * static_cast<long long>(blockIdx.x)] = smid;
                p.trace[1024 + 4 * static_cast<long long>(blockIdx.x) + 1] = static_cast<long long>(globaltimer_ns());
            }
        }
    }
    // (batch, head) major, key block minor: the 16 CTAs that share one (batch, head) run close together in time, so its
    // Q / dO blocks and its dQ accumulator tiles stay in L2.  (Key-block-major order made the kernel DRAM-bound: ncu
    // measured 3.2 GB of traffic per launch — every Q / dO re-read and every dQ reduce-add went to HBM.)
    const int bh = blockIdx.x / nblk;
    const int j = blockIdx.x - bh * nblk;                // heavy (small j) first inside a (batch, head)
    const int head = bh % p.heads, b = bh / p.heads;
    const int row_base = b * T;
    const int i0 = CAUSAL ? j : 0;
    const int n_it = nblk - i0;
    const long long vec_base = (static_cast<long long>(head) * p.batch + b) * p.t_pad;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023) { printf("kx attn_bwd: dynamic smem base not 1024-aligned\n"); __trap(); }
        mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_pds, BW_COMPUTE); mbar_init(bar_dq, 1); mbar_init(bar_dqr, 128); mbar_init(bar_fin, 1); mbar_init(bar_sfree, BW_COMPUTE);
        fence_mbar_init();
    }
    if (warp == BW_W_TMA) {
        // The producer's own barriers are initialised by the producer warp, and K_j, V_j and the first query blocks are requested
        // BEFORE the TMEM allocation and the CTA-wide barrier: a CTA lives ~20 us and every CTA pays this prologue (the per-CTA
        // records of the trace build: 2.3 us from entry to the first ready score tile, 11 % of the kernel).
        if (lane == 0) {
            mbar_init(kv_full, 1);
            for (int s = 0; s < BW_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            fence_mbar_init();
            mbar_arrive_expect_tx(kv_full, 2 * BW_TILE);
            tma_load_2d(&tmK, kv_full, smem + BW_SMEM_K, head * 64, row_base + j * 128, kEvictFirst);
            tma_load_2d(&tmV, kv_full, smem + BW_SMEM_V, head * 64, row_base + j * 128, kEvictFirst);
            for (int it = 0; it < BW_SLOTS && it < n_it; ++it) {
                const int i = i0 + it;
                mbar_arrive_expect_tx(&full[it], 2 * BW_TILE + 1024);
                tma_load_2d(&tmQ, &full[it], smem + BW_SMEM_Q + it * BW_TILE, head * 64, row_base + i * 128, kEvictLast);
                tma_load_2d(&tmDO, &full[it], smem + BW_SMEM_DO + it * BW_TILE, head * 64, row_base + i * 128, kEvictLast);
                bulk_load_1d(s_ld + it * 64, p.nld + vec_base + i * 128, 1024, &full[it]);
            }
            prefetch_tmap(&tmDQ);
        }
        __syncwarp();
        tmem_alloc<1>(tmem_slot, BW_TMEM_COLS);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= BW_W_DRAIN) {
        // ================= dQ drain: thread = query row r =================
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int sw = r & 7;
        const bool issuer = (threadIdx.x == BW_W_DRAIN * 32);
        for (int it = 0; it < n_it; ++it) {
            const int i = i0 + it;
            mbar_wait_lean(bar_dq, it & 1);
            tc_fence_after();
            if (issuer) tma_store_wait_read<0>();                    // the previous reduction has read the staging tiles
            named_bar_sync(2, 128);
#pragma unroll
            for (int h = 0; h < 4; ++h) {                            // 16 columns at a time
                uint32_t qv[16];
                tmem_ld16(tmem_base + lane_addr + BW_T_DQ + h * 16, qv);
                tmem_ld_wait();
                uint8_t* dq_row = smem + BW_SMEM_DQ + (h >> 1) * BW_TILE + r * 128;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<float4*>(dq_row + ((((h & 1) * 4 + u) ^ sw) << 4)) =
                        make_float4(__uint_as_float(qv[4 * u]) * p.scale, __uint_as_float(qv[4 * u + 1]) * p.scale,
                                    __uint_as_float(qv[4 * u + 2]) * p.scale, __uint_as_float(qv[4 * u + 3]) * p.scale);
            }
            tc_fence_before();
            mbar_arrive(bar_dqr);                                    // the next dQ MMA may overwrite the accumulator
            fence_proxy_async_smem();
            named_bar_sync(3, 128);
            // rows of the tile beyond T are exact zeros (their dS is masked); rows beyond the matrix are clipped by TMA
            if (issuer) {
                tma_reduce_add_2d(&tmDQ, smem + BW_SMEM_DQ, head * 64, row_base + i * 128);
                tma_reduce_add_2d(&tmDQ, smem + BW_SMEM_DQ + BW_TILE, head * 64 + 32, row_base + i * 128);
                tma_store_commit();
            }
        }
        if (issuer) tma_store_wait<0>();
    } else if (warp >= BW_W_TMA) {
        if (warp == BW_W_TMA) {
            if (elect_one()) {
                // ================= TMA producer (K_j, V_j and the first BW_SLOTS query blocks were requested in the prologue) =================
                for (int it = BW_SLOTS; it < n_it; ++it) {
                    const int i = i0 + it, s = it % BW_SLOTS;
                    mbar_wait_lean(&empty[s], ((it / BW_SLOTS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full[s], 2 * BW_TILE + 1024);
                    tma_load_2d(&tmQ, &full[s], smem + BW_SMEM_Q + s * BW_TILE, head * 64, row_base + i * 128, kEvictLast);
                    tma_load_2d(&tmDO, &full[s], smem + BW_SMEM_DO + s * BW_TILE, head * 64, row_base + i * 128, kEvictLast);
                    bulk_load_1d(s_ld + s * 64, p.nld + vec_base + i * 128, 1024, &full[s]);
                }
            }
        } else if (warp == BW_W_MMA) {
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
                // Straight-line issue (a rolled loop made every UTCHMMA cost ~90 cycles on a scheduler shared with five busy
                // warps); descriptors travel as (low, high) words so the unrolled steps need no 64-bit register pairs.
                auto lo = [](uint64_t d) { return static_cast<uint32_t>(d); };
                auto hi = [](uint64_t d) { return static_cast<uint32_t>(d >> 32); };
                auto mma_ss = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t astep, uint32_t bstep, int steps, uint32_t idesc, uint32_t acc0) {
                    const uint32_t al = lo(a), ah = hi(a), bl = lo(b), bh = hi(b);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < steps) umma_bf16_lohi(d, al + k * astep, ah, bl + k * bstep, bh, idesc, (k > 0) ? 1u : acc0);
                };
                auto mma_ts = [&](uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t acc0) {
                    const uint32_t bl = lo(b), bh = hi(b);
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_bf16_ts_lohi(d, a_tmem + k * 8, bl + k * 128, bh, idesc_kv, (k > 0) ? 1u : acc0);
                };
                const uint32_t slot_step = BW_TILE >> 4;           // descriptor address units (16 B) between ring slots
                const uint64_t q_kmaj0 = make_desc_sw128(smem_u32(smem + BW_SMEM_Q));
                const uint64_t do_kmaj0 = make_desc_sw128(smem_u32(smem + BW_SMEM_DO));
                const uint64_t q_mn0 = make_desc_sw128(smem_u32(smem + BW_SMEM_Q), BW_TILE);
                const uint64_t do_mn0 = make_desc_sw128(smem_u32(smem + BW_SMEM_DO), BW_TILE);
                mbar_wait_lean(kv_full, 0);
                mbar_wait_lean(&full[0], 0);
                tc_fence_after();
                mma_ss(t_s, k_kmaj, q_kmaj0, 2, 2, 4, idesc_s, 0u);
                umma_commit(bar_s);
                mma_ss(t_dp, v_kmaj, do_kmaj0, 2, 2, 4, idesc_s, 0u);
                umma_commit(bar_dp);
                const uint32_t t_p = tmem_base + BW_T_P;
                const uint64_t ds_kmaj = make_desc_sw128(smem_u32(smem + BW_SMEM_DS));       // dS^T as a K-major A operand (rows = keys)
                for (int it = 0; it < n_it; ++it) {
                    const int s = it % BW_SLOTS, sn = (it + 1) % BW_SLOTS;
                    const bool has_next = it + 1 < n_it;
                    const uint32_t acc = it > 0 ? 1u : 0u;
                    KX_BT(1, it, 0);
                    if (has_next) {                      // scores of the next block as soon as this block's are in registers
                        mbar_wait_lean(bar_sfree, it & 1);
                        mbar_wait_lean(&full[sn], ((it + 1) / BW_SLOTS) & 1);
                        tc_fence_after();
                        mma_ss(t_s, k_kmaj, q_kmaj0 + sn * slot_step, 2, 2, 4, idesc_s, 0u);
                        umma_commit(bar_s);
                        mma_ss(t_dp, v_kmaj, do_kmaj0 + sn * slot_step, 2, 2, 4, idesc_s, 0u);
                        umma_commit(bar_dp);
                    }
                    KX_BT(1, it, 1);
                    mbar_wait_lean(bar_pds, it & 1);
                    tc_fence_after();
                    KX_BT(1, it, 2);
                    mma_ts(t_dv, t_p, do_mn0 + s * slot_step, acc);          // dV: 16 queries per step = 8 TMEM columns of bf16 pairs, 16 dO rows = 2 KB
                    {                                                         // dK: A = dS^T from smem, K-major: 4 steps of 32 B in each 64-query atom
                        const uint32_t al = lo(ds_kmaj), ah = hi(ds_kmaj);
                        const uint64_t qm = q_mn0 + s * slot_step;
                        const uint32_t bl = lo(qm), bh = hi(qm);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16_lohi(t_dk, al + (k >> 2) * slot_step + (k & 3) * 2, ah, bl + k * 128, bh, idesc_kv, (k > 0) ? 1u : acc);
                    }
                    KX_BT(1, it, 3);
                    if (it > 0) {                        // dQ of the previous iteration has been read out of TMEM
                        mbar_wait_lean(bar_dqr, (it - 1) & 1);
                        tc_fence_after();
                    }
                    KX_BT(1, it, 4);
                    mma_ss(t_dq, ds_mn, k_mn, 128, 128, 8, idesc_dq, 0u);   // dQ: 16 keys per step = 16 rows of dS^T (smem) / K = 2 KB
                    umma_commit(bar_dq);                 // also: P^T and the smem copy of dS^T may be overwritten
                    umma_commit(&empty[s]);
                    if (!has_next) umma_commit(bar_fin);
                    KX_BT(1, it, 5);
                }
            }
        }
    } else {
        // ================= compute: thread = key row r (TMEM lane), warpgroup g = query column quarter =================
        const int g = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int kg = j * 128 + r;
        const uint64_t sl2x2 = pack_f32x2(p.scale_log2, p.scale_log2);
        const int sw = r & 7;
        // dS^T staging: queries [32g, 32g+32) = chunks 4*(g&1) .. +3 of row r in MN atom g>>1
        uint8_t* ds_row = smem + BW_SMEM_DS + (g >> 1) * BW_TILE + r * 128;
        const int ch0 = (g & 1) * 4;

        // DROP: keep bits of queries 32g .. 32g+31 for this key, loaded one query block ahead (the round trip to L2 / DRAM at
        // the top of every iteration sat on the critical path of the compute warps)
        auto mask_word = [&](int i) {
            return __ldg(p.drop_mask + ((((static_cast<long long>(bh) * nblk + i) * nblk + j) * 4 + g) << 7) + r);
        };
        uint32_t mw_next = 0xffffffffu;
        if constexpr (DROP) mw_next = mask_word(i0);
        for (int it = 0; it < n_it; ++it) {
            const int i = i0 + it, s = it % BW_SLOTS;
            if (threadIdx.x == 0) KX_BT(0, it, 0);
            const uint32_t mw = mw_next;
            if constexpr (DROP) {
                if (it + 1 < n_it) mw_next = mask_word(i + 1);
            }
            mbar_wait_lean(&full[s], (it / BW_SLOTS) & 1);                  // (-lse, -delta) of this query block are in smem
            mbar_wait_lean(bar_s, it & 1);
            tc_fence_after();
            if (threadIdx.x == 0) KX_BT(0, it, 1);
            if constexpr (TRACE) {
                if (p.trace != nullptr && threadIdx.x == 0 && it == 0)
                    p.trace[1024 + 4 * static_cast<long long>(blockIdx.x) + 2] = static_cast<long long>(globaltimer_ns());
            }
            const float4* ld = s_ld + s * 64 + g * 16;          // {-lse(q0), -lse(q1), -delta(q0), -delta(q1)} per query pair
            const bool edge = (CAUSAL && i == j) || (i * 128 + 128 > T) || (j * 128 + 128 > T);
            mbar_wait_lean(bar_dp, it & 1);
            tc_fence_after();
            if (threadIdx.x == 0) KX_BT(0, it, 2);
            uint32_t pp[16], dd[16];
            // P^T = exp2(S^T * scale * log2e - lse), dS^T = P^T * (dP^T - delta) in fp32, both packed to bf16; masked entries
            // of the diagonal / tail blocks become exact zeros.  Two 16-column halves: the live set has to fit the 80
            // registers that 768 threads leave per thread.
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t sv[16], dpv[16];
                tmem_ld16(tmem_base + lane_addr + BW_T_S + g * 32 + hf * 16, sv);
                tmem_ld16(tmem_base + lane_addr + BW_T_DP + g * 32 + hf * 16, dpv);
                tmem_ld_wait();
                if (hf == 1) {                                    // S^T and dP^T of this block are in registers everywhere soon
                    tc_fence_before();
                    mbar_arrive(bar_sfree);
                }
                if (!edge) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 l = ld[hf * 8 + c];
                        float x0, x1;
                        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1])), sl2x2, pack_f32x2(l.x, l.y)), x0, x1);
                        float e0, e1;
                        // 3 of 8 pairs on the FMA pipes (the loop is MUFU-bound otherwise); none with dropout, whose mask logic already
                        // loads the FMA / ALU pipes (A/B on one box: 721 us with {1, 4, 6}, 703 us with two pairs, 683 us with none;
                        // without dropout 614 us either way)
                        if (!DROP && (c == 1 || c == 4 || c == 6)) {
                            exp2_poly_x2(x0, x1, e0, e1);
                        } else {
                            e0 = ex2_approx(x0);
                            e1 = ex2_approx(x1);
                        }
                        float dp0 = __uint_as_float(dpv[2 * c]), dp1 = __uint_as_float(dpv[2 * c + 1]);
                        uint32_t pmask = 0xffffffffu;
                        if constexpr (DROP) {              // dS = P o (M o dP / keep - delta); the dV operand is M o P (1/keep on the way out)
                            const uint32_t b2 = (mw >> (hf * 16 + 2 * c)) & 3u;
                            dp0 = (b2 & 1u) ? dp0 * p.inv_keep : 0.f;
                            dp1 = (b2 & 2u) ? dp1 * p.inv_keep : 0.f;
                            pmask = ((b2 & 1u) ? 0x0000ffffu : 0u) | ((b2 & 2u) ? 0xffff0000u : 0u);
                        }
                        const uint64_t t = fadd2(pack_f32x2(dp0, dp1), pack_f32x2(l.z, l.w));
                        float d0, d1;
                        unpack_f32x2(fmul2(pack_f32x2(e0, e1), t), d0, d1);
                        pp[hf * 8 + c] = pack_bf16(e0, e1) & pmask;
                        dd[hf * 8 + c] = pack_bf16(d0, d1);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 l = ld[hf * 8 + c];
                        const int q0 = i * 128 + g * 32 + hf * 16 + 2 * c;
                        const bool ok0 = kg < T && q0 < T && (!CAUSAL || q0 >= kg);
                        const bool ok1 = kg < T && q0 + 1 < T && (!CAUSAL || q0 + 1 >= kg);
                        const float p0 = ok0 ? ex2_approx(fmaf(__uint_as_float(sv[2 * c]), p.scale_log2, l.x)) : 0.f;
                        const float p1 = ok1 ? ex2_approx(fmaf(__uint_as_float(sv[2 * c + 1]), p.scale_log2, l.y)) : 0.f;
                        float dp0 = __uint_as_float(dpv[2 * c]), dp1 = __uint_as_float(dpv[2 * c + 1]);
                        uint32_t pmask = 0xffffffffu;
                        if constexpr (DROP) {
                            const uint32_t b2 = (mw >> (hf * 16 + 2 * c)) & 3u;
                            dp0 = (b2 & 1u) ? dp0 * p.inv_keep : 0.f;
                            dp1 = (b2 & 2u) ? dp1 * p.inv_keep : 0.f;
                            pmask = ((b2 & 1u) ? 0x0000ffffu : 0u) | ((b2 & 2u) ? 0xffff0000u : 0u);
                        }
                        // rows beyond T may hold inf / nan in dP or delta: never form 0 * nan
                        const float d0 = ok0 ? p0 * (dp0 + l.z) : 0.f;
                        const float d1 = ok1 ? p1 * (dp1 + l.w) : 0.f;
                        pp[hf * 8 + c] = pack_bf16(p0, p1) & pmask;
                        dd[hf * 8 + c] = pack_bf16(d0, d1);
                    }
                }
            }
            if (threadIdx.x == 0) KX_BT(0, it, 4);
            if (it > 0) mbar_wait_lean(bar_dq, (it - 1) & 1);               // dV_{i-1} / dK_{i-1} / dQ_{i-1} have consumed P^T and the smem dS^T
            if (threadIdx.x == 0) KX_BT(0, it, 5);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)                                   // dS^T row -> smem (A of the dK and dQ MMAs), SWIZZLE_128B
                *reinterpret_cast<uint4*>(ds_row + (((ch0 + ch) ^ sw) << 4)) = make_uint4(dd[4 * ch], dd[4 * ch + 1], dd[4 * ch + 2], dd[4 * ch + 3]);
            fence_proxy_async_smem();
            tmem_st16(tmem_base + lane_addr + BW_T_P + g * 16, pp);         // P^T (bf16 pairs), its own TMEM region
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_pds);
            if (threadIdx.x == 0) KX_BT(0, it, 6);
        }
        // ---- epilogue: dV_j, dK_j
        mbar_wait_lean(bar_fin, 0);
        tc_fence_after();
        uint32_t vv[16], kk[16];
        tmem_ld16(tmem_base + lane_addr + BW_T_DV + g * 16, vv);
        tmem_ld16(tmem_base + lane_addr + BW_T_DK + g * 16, kk);
        tmem_ld_wait();
        if (kg < T) {
            const long long off = static_cast<long long>(row_base + kg) * p.ld_dkv + head * 64 + g * 16;
#pragma unroll
            for (int u = 0; u < 16; ++u) kk[u] = __float_as_uint(__uint_as_float(kk[u]) * p.scale);
            if (p.k_cos != nullptr) {
                // dK = R^T(dK): the transpose of the xPos rotation the QKV epilogue applied to k (pairs (2u, 2u + 1) of this
                // thread's 16 columns; previously a second pass over the stored bf16 dK in attn_bwd_finish_kernel)
                const float4* cp = reinterpret_cast<const float4*>(p.k_cos + static_cast<long long>(kg) * 32 + g * 8);
                const float4* sp = reinterpret_cast<const float4*>(p.k_sin + static_cast<long long>(kg) * 32 + g * 8);
                const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float a = __uint_as_float(kk[2 * u]), b = __uint_as_float(kk[2 * u + 1]);
                    kk[2 * u] = __float_as_uint(a * cc[u] + b * ss[u]);
                    kk[2 * u + 1] = __float_as_uint(b * cc[u] - a * ss[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                uint4 a, c;
                const float vs = DROP ? p.inv_keep : 1.0f;          // dV accumulated (M o P)^T dO: the 1/keep factor goes on here
                a.x = pack_bf16(__uint_as_float(vv[8 * u]) * vs, __uint_as_float(vv[8 * u + 1]) * vs);
                a.y = pack_bf16(__uint_as_float(vv[8 * u + 2]) * vs, __uint_as_float(vv[8 * u + 3]) * vs);
                a.z = pack_bf16(__uint_as_float(vv[8 * u + 4]) * vs, __uint_as_float(vv[8 * u + 5]) * vs);
                a.w = pack_bf16(__uint_as_float(vv[8 * u + 6]) * vs, __uint_as_float(vv[8 * u + 7]) * vs);
                c.x = pack_bf16(__uint_as_float(kk[8 * u]), __uint_as_float(kk[8 * u + 1]));
                c.y = pack_bf16(__uint_as_float(kk[8 * u + 2]), __uint_as_float(kk[8 * u + 3]));
                c.z = pack_bf16(__uint_as_float(kk[8 * u + 4]), __uint_as_float(kk[8 * u + 5]));
                c.w = pack_bf16(__uint_as_float(kk[8 * u + 6]), __uint_as_float(kk[8 * u + 7]));
                *reinterpret_cast<uint4*>(p.dv + off + 8 * u) = a;
                *reinterpret_cast<uint4*>(p.dk + off + 8 * u) = c;
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == BW_W_TMA) tmem_dealloc<1>(tmem_base, BW_TMEM_COLS);
    if constexpr (TRACE) {
        if (p.trace != nullptr && threadIdx.x == 0)
            p.trace[1024 + 4 * static_cast<long long>(blockIdx.x) + 3] = static_cast<long long>(globaltimer_ns());
    }
}

// nld[h][b][t/2] = {-lse(t0), -lse(t1), -delta(t0), -delta(t1)}, delta = sum_d dO[b,t,h,d] * O[b,t,h,d]  (8 lanes per (row, head)).
// Both come out negated so that the main kernel's packed FFMA2 / FADD2 take them as plain addends.
// The same pass zeroes the fp32 dQ accumulator (the lane that reads 16 bytes of (row, head) clears the matching 32 bytes): it
// replaces a separate 268 MB memset at the C3 shape.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ld_o, const __nv_bfloat16* __restrict__ d_o, long long ld_do,
                  const float* __restrict__ lse, float2* __restrict__ nld, float* __restrict__ dq_zero, int batch, int heads,
                  int seq_len, int t_pad) {
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
    if (active) {
        float4* z = reinterpret_cast<float4*>(dq_zero + (row * heads + head) * 64 + sub * 8);
        z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
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
        const long long o_ = (static_cast<long long>(head) * batch + b) * t_pad + t;
        float* dst = reinterpret_cast<float*>(nld) + ((o_ >> 1) << 2) + (o_ & 1);      // per query pair: {-lse0, -lse1, -delta0, -delta1}
        dst[0] = -lse[o_];
        dst[2] = -s;
    }
}

// dq (bf16) = R^T(dq_accum), R = the xPos rotation of the QKV epilogue on q (dK gets its rotation in the main kernel's
// epilogue).  ld in elements of the bf16 matrix.
__global__ void __launch_bounds__(256)
attn_bwd_finish_kernel(const float* __restrict__ dq_accum, __nv_bfloat16* __restrict__ dq, long long ld, int rows, int d_model,
                       int seq_len, const float* __restrict__ q_cos, const float* __restrict__ q_sin) {
    const int vec_per_row = d_model >> 3;
    const long long total = static_cast<long long>(rows) * vec_per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / vec_per_row);
        const int col = static_cast<int>(i - static_cast<long long>(row) * vec_per_row) * 8;
        const float4 a = *reinterpret_cast<const float4*>(dq_accum + static_cast<long long>(row) * d_model + col);
        const float4 c = *reinterpret_cast<const float4*>(dq_accum + static_cast<long long>(row) * d_model + col + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        float o[8];
        if (q_cos != nullptr) {
            const int t = row % seq_len;
            const int j0 = (col & 63) >> 1;
            const float4 cv = __ldg(reinterpret_cast<const float4*>(q_cos + t * 32 + j0));
            const float4 sv = __ldg(reinterpret_cast<const float4*>(q_sin + t * 32 + j0));
            const float cc[4] = {cv.x, cv.y, cv.z, cv.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w};
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
        *reinterpret_cast<uint4*>(dq + static_cast<long long>(row) * ld + col) = q;
    }
}

}  // namespace kx

using namespace kx;

// Profiling aid for kx_attn_bwd (causal): with a device buffer of 2*32*16 + 4 * (number of CTAs) int64 installed, one CTA of every
// launch (the one named by word 1023 of the buffer; negative = CTA 0) records clock64 stamps [role: compute thread 0, MMA thread]
// [iteration][point], and every CTA b leaves {SM id, globaltimer ns at entry, at its first ready score tile, at exit} at
// word 1024 + 4 b; NULL = off.
extern "C" int kx_attn_bwd_set_trace(long long* device_buffer) {
    g_attn_bwd_trace = device_buffer;
    return KX_OK;
}

static int attn_bwd_impl(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                         const void* d_out, long long ld_dout, const float* lse, void* dq, void* dk, void* dv,
                         long long ld_dqkv, float* dq_accum, float* delta, const float* xq_cos, const float* xq_sin,
                         const float* xk_cos, const float* xk_sin, int batch, int heads, int seq_len, int causal, float scale,
                         float drop_p, const unsigned int* drop_mask, cudaStream_t stream) {
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
    CUtensorMap tq, tk, tv, tdo, tdq;
    if (!make_tmap_f32_2d(&tdq, dq_accum, (uint64_t)heads * 64, rows, (uint64_t)heads * 64 * 4, 32, 128)) return KX_ERR_TMAP;
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
    {
        const long long total = static_cast<long long>(rows) * heads * 8;
        attn_delta_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(out), ld_out, reinterpret_cast<const __nv_bfloat16*>(d_out), ld_dout, lse,
            reinterpret_cast<float2*>(delta), dq_accum, batch, heads, seq_len, t_pad);
        int st = check_launch("kx_attn_bwd (delta)");
        if (st != KX_OK) return st;
    }
    AttnBwdParams p;
    p.nld = reinterpret_cast<const float2*>(delta); p.dq_accum = dq_accum;
    p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.ld_dkv = ld_dqkv;
    p.seq_len = seq_len; p.heads = heads; p.batch = batch; p.t_pad = t_pad;
    p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
    p.trace = g_attn_bwd_trace;
    p.drop_mask = drop_mask;
    p.inv_keep = 1.0f;
    p.k_cos = xk_cos; p.k_sin = xk_sin;
    if (drop_mask != nullptr) {
        if (!causal || !(drop_p > 0.f && drop_p < 1.f) || (reinterpret_cast<uintptr_t>(drop_mask) & 15)) {
            set_error("kx_attn_bwd_dropout: needs causal attention, 0 < p < 1 and the 16-byte aligned key_mask of kx_attn_dropout_masks");
            return KX_ERR_ARG;
        }
        const unsigned thr12 = static_cast<unsigned>((1.0 - static_cast<double>(drop_p)) * 4096.0 + 0.5);   // as kx_attn_dropout_masks
        p.inv_keep = 4096.0f / static_cast<float>(thr12);
    }
    const long long ctas = static_cast<long long>(nblk) * heads * batch;
    if (ctas > 0x7fffffffLL) { set_error("kx_attn_bwd: too many tiles"); return KX_ERR_ARG; }
    if (causal && p.trace != nullptr) {           // profiling aid (kx_attn_bwd_set_trace): same kernel with clock64 stamps
        static bool tattr = false;
        if (!tattr) { cudaFuncSetAttribute(attn_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES); tattr = true; }
        attn_bwd_kernel<true, true><<<static_cast<unsigned>(ctas), BW_THREADS, BW_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, tdq, p);
    } else if (causal && drop_mask != nullptr) {
        static bool dattr = false;
        if (!dattr) { cudaFuncSetAttribute(attn_bwd_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES); dattr = true; }
        attn_bwd_kernel<true, false, true><<<static_cast<unsigned>(ctas), BW_THREADS, BW_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, tdq, p);
    } else if (causal) attn_bwd_kernel<true><<<static_cast<unsigned>(ctas), BW_THREADS, BW_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, tdq, p);
    else attn_bwd_kernel<false><<<static_cast<unsigned>(ctas), BW_THREADS, BW_SMEM_BYTES, stream>>>(tq, tk, tv, tdo, tdq, p);
    int st = check_launch("kx_attn_bwd");
    if (st != KX_OK) return st;
    {
        const long long total = static_cast<long long>(rows) * (d_model / 8);
        const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sms) * 8));
        attn_bwd_finish_kernel<<<blocks, 256, 0, stream>>>(dq_accum, reinterpret_cast<__nv_bfloat16*>(dq), ld_dqkv, static_cast<int>(rows),
                                                          static_cast<int>(d_model), seq_len, xq_cos, xq_sin);
        st = check_launch("kx_attn_bwd (finish)");
    }
    return st;
}

extern "C" int kx_attn_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                           const void* d_out, long long ld_dout, const float* lse, void* dq, void* dk, void* dv,
                           long long ld_dqkv, float* dq_accum, float* delta, const float* xq_cos, const float* xq_sin,
                           const float* xk_cos, const float* xk_sin, int batch, int heads, int seq_len, int causal, float scale,
                           cudaStream_t stream) {
    return attn_bwd_impl(q, k, v, ld_qkv, out, ld_out, d_out, ld_dout, lse, dq, dk, dv, ld_dqkv, dq_accum, delta, xq_cos, xq_sin,
                         xk_cos, xk_sin, batch, heads, seq_len, causal, scale, 0.f, nullptr, stream);
}

extern "C" int kx_attn_bwd_dropout(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                                   const void* d_out, long long ld_dout, const float* lse, void* dq, void* dk, void* dv,
                                   long long ld_dqkv, float* dq_accum, float* delta, const float* xq_cos, const float* xq_sin,
                                   const float* xk_cos, const float* xk_sin, int batch, int heads, int seq_len, int causal,
                                   float scale, float drop_p, const unsigned int* drop_mask, cudaStream_t stream) {
    if (!drop_mask) { set_error("kx_attn_bwd_dropout: null mask"); return KX_ERR_ARG; }
    return attn_bwd_impl(q, k, v, ld_qkv, out, ld_out, d_out, ld_dout, lse, dq, dk, dv, ld_dqkv, dq_accum, delta, xq_cos, xq_sin,
                         xk_cos, xk_sin, batch, heads, seq_len, causal, scale, drop_p, drop_mask, stream);
}
