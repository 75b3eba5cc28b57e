// kx_attn_fwd: flash-attention forward, head_dim 64, on tcgen05 tensor cores.
//
// Replaces the materialised-score attention of the reference's dependencies:
//   decoder  (causal)     torchscale MultiheadAttention: bmm -> nan_to_num -> +triu(-inf) ->
//                         softmax(fp32) -> bmm -> head merge            
//   ViT      (non-causal) HF CLIPAttention eager/sdpa core              ([HF] modeling_clip.py:318-331, k4)
//
// One CTA = one 128-row query tile of one (batch, head); 2 CTAs co-reside per SM.
//   warps 0-3  softmax (thread == query row): tcgen05.ld S from TMEM, online softmax in fp32,
//              P (bf16) -> swizzled smem, running O in registers (O = O*alpha + P.V per block)
//   warp 4     TMA producer: Q once, K/V 128-row blocks through 2-slot rings
//   warp 5     MMA issuer: S = Q.K^T (M128 N128 K64) and PV = P.V (M128 N64 K128) into TMEM
// Scores never touch HBM; the causal mask is a predicate on the diagonal block only and KV
// blocks above the diagonal are skipped.
#include "kx_internal.h"
#include "ptx.cuh"
#include <cstdlib>

namespace kx {

constexpr int AT_BM = 128;          // query rows per CTA
constexpr int AT_BN = 128;          // keys per block
constexpr int AT_D = 64;            // head dim
constexpr int AT_THREADS = 192;
constexpr int AT_TILE_BYTES = 128 * 64 * 2;   // 16 KB: one [128 x 64] bf16 tile (Q, K or V block)
constexpr int AT_SMEM_Q = 0;
constexpr int AT_SMEM_K = AT_SMEM_Q + AT_TILE_BYTES;          // 2 slots
constexpr int AT_SMEM_V = AT_SMEM_K + 2 * AT_TILE_BYTES;      // 2 slots
constexpr int AT_SMEM_P = AT_SMEM_V + 2 * AT_TILE_BYTES;      // [2 k-atoms][128 rows][128 B]
constexpr int AT_SMEM_BAR = AT_SMEM_P + 2 * AT_TILE_BYTES;
constexpr int AT_SMEM_BYTES = AT_SMEM_BAR + 128;
constexpr int AT_TMEM_COLS = 256;   // S: [0,128)  PV: [128,192)

struct AttnParams {
    __nv_bfloat16* out;
    long long ld_out;
    int seq_len, heads, num_q_tiles;
    float scale_log2;               // scale * log2(e)
};

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;
    uint64_t* s_empty = bars + 10;
    uint64_t* p_full = bars + 11;
    uint64_t* o_full = bars + 12;
    uint64_t* o_empty = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = p.num_q_tiles - 1 - blockIdx.x;     // heaviest (longest KV range) tiles first
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int T = p.seq_len;
    const int q0 = qt * AT_BM;
    const int n_blocks = CAUSAL ? (qt + 1) : (T + AT_BN - 1) / AT_BN;
    const int row_base = b * T;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023) { printf("kx attn: dynamic smem base not 1024-aligned\n"); __trap(); }
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        }
        mbar_init(s_full, 1); mbar_init(s_empty, 128);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1); mbar_init(o_empty, 128);
        fence_mbar_init();
    }
    if (warp == 4) {
        if (lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); }
        tmem_alloc<1>(tmem_slot, AT_TMEM_COLS);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + 128;

    if (warp == 4) {
        if (lane == 0) {
            // ================= TMA producer =================
            mbar_arrive_expect_tx(q_full, AT_TILE_BYTES);
            tma_load_2d(&tmQ, q_full, smem + AT_SMEM_Q, head * AT_D, row_base + q0, kEvictFirst);
            for (int j = 0; j < n_blocks; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[s], AT_TILE_BYTES);
                tma_load_2d(&tmK, &k_full[s], smem + AT_SMEM_K + s * AT_TILE_BYTES, head * AT_D, row_base + j * AT_BN, kEvictLast);
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[s], AT_TILE_BYTES);
                tma_load_2d(&tmV, &v_full[s], smem + AT_SMEM_V + s * AT_TILE_BYTES, head * AT_D, row_base + j * AT_BN, kEvictLast);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc_s = make_idesc_bf16(128, AT_BN, 0, 0);   // Q (K-major) x K (K-major)
            constexpr uint32_t idesc_o = make_idesc_bf16(128, AT_D, 0, 1);    // P (K-major) x V (MN-major)
            const uint64_t qdesc = make_desc_sw128(smem_u32(smem + AT_SMEM_Q));
            const uint64_t pdesc = make_desc_sw128(smem_u32(smem + AT_SMEM_P));
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint64_t kdesc = make_desc_sw128(smem_u32(smem + AT_SMEM_K + s * AT_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k) umma_bf16<1>(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
                umma_commit(&k_empty[s]);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_blocks; ++j) {
                if (j + 1 < n_blocks) {
                    mbar_wait(s_empty, j & 1);          // softmax has pulled S(j) out of TMEM
                    tc_fence_after();
                    issue_s(j + 1);
                }
                const int s = j & 1;
                mbar_wait(p_full, j & 1);               // P(j) is in smem
                mbar_wait(&v_full[s], (j >> 1) & 1);
                if (j > 0) mbar_wait(o_empty, (j - 1) & 1);   // PV(j-1) has been read out of TMEM
                tc_fence_after();
                const uint64_t vdesc = make_desc_sw128(smem_u32(smem + AT_SMEM_V + s * AT_TILE_BYTES), AT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < AT_BN / 16; ++k) {
                    // A = P: k-atom (k/4) of 16 KB, 32 B per K step inside it.  B = V (MN-major): 16 kv rows = 2 KB per K step.
                    const uint64_t ad = pdesc + (k >> 2) * (AT_TILE_BYTES >> 4) + 2 * (k & 3);
                    const uint64_t bd = vdesc + k * (2048 >> 4);
                    umma_bf16<1>(tmem_o, ad, bd, idesc_o, k != 0);
                }
                umma_commit(&v_empty[s]);
                umma_commit(o_full);
            }
        }
    } else {
        // ================= softmax / output (warps 0-3, thread == query row) =================
        const int r = warp * 32 + lane;                 // row inside the tile == TMEM lane
        const int qrow = q0 + r;
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        float o_acc[AT_D];
#pragma unroll
        for (int i = 0; i < AT_D; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
        uint8_t* p_row = smem + AT_SMEM_P + r * 128;
        const int sw = r & 7;

        for (int j = 0; j < n_blocks; ++j) {
            const int kv0 = j * AT_BN;
            const bool need_mask = (j == n_blocks - 1);      // diagonal block (causal) / ragged tail
            // valid keys for this row in this block: c < limit
            int limit = AT_BN;
            if (need_mask) {
                limit = T - kv0;
                if (CAUSAL) limit = min(limit, qrow - kv0 + 1);
                limit = max(limit, 0);
            }
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // ---- pass 1: row max
            float m_blk = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < AT_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_s + lane_addr + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float x = __uint_as_float(v[i]);
                    if (!need_mask || (c * 32 + i) < limit) m_blk = fmaxf(m_blk, x);
                }
            }
            const float m_new = fmaxf(m_run, m_blk);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = exp2f((m_run - m_use) * p.scale_log2);   // m_run = -inf -> 0
            const float m_scaled = m_use * p.scale_log2;

            // ---- fold in PV(j-1) (it was computed against the previous running max)
            if (j > 0) {
                mbar_wait(o_full, (j - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < AT_D / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_o + lane_addr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = o_acc[c * 32 + i] * alpha_prev + __uint_as_float(v[i]);
                }
                tc_fence_before();
                mbar_arrive(o_empty);
            }

            // ---- pass 2: p = exp2(s*scale - m), row sum, P -> smem (bf16, 128B-swizzled K-major)
            float l_blk = 0.f;
#pragma unroll 1
            for (int c = 0; c < AT_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_s + lane_addr + c * 32, v);
                tmem_ld_wait();
                float e[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float x = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_scaled));
                    if (need_mask && (c * 32 + i) >= limit) x = 0.f;
                    e[i] = x;
                    l_blk += x;
                }
                // columns [32c, 32c+32) = 64 B = chunks (4*(c&1) .. +3) of k-atom (c >> 1)
                uint8_t* atom = p_row + (c >> 1) * AT_TILE_BYTES;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 q;
                    q.x = pack_bf16(e[8 * g + 0], e[8 * g + 1]);
                    q.y = pack_bf16(e[8 * g + 2], e[8 * g + 3]);
                    q.z = pack_bf16(e[8 * g + 4], e[8 * g + 5]);
                    q.w = pack_bf16(e[8 * g + 6], e[8 * g + 7]);
                    const int chunk = (4 * (c & 1) + g) ^ sw;
                    *reinterpret_cast<uint4*>(atom + chunk * 16) = q;
                }
            }
            l_run = l_run * alpha + l_blk;
            m_run = m_new;
            alpha_prev = alpha;
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(s_empty);
            mbar_arrive(p_full);
        }
        // ---- last PV block
        {
            const int j = n_blocks - 1;
            mbar_wait(o_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < AT_D / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_o + lane_addr + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = o_acc[c * 32 + i] * alpha_prev + __uint_as_float(v[i]);
            }
            tc_fence_before();
        }
        if (qrow < T) {
            const float inv_l = 1.0f / l_run;
            __nv_bfloat16* o = p.out + static_cast<long long>(row_base + qrow) * p.ld_out + head * AT_D;
#pragma unroll
            for (int g = 0; g < AT_D / 8; ++g) {
                uint4 q;
                q.x = pack_bf16(o_acc[8 * g + 0] * inv_l, o_acc[8 * g + 1] * inv_l);
                q.y = pack_bf16(o_acc[8 * g + 2] * inv_l, o_acc[8 * g + 3] * inv_l);
                q.z = pack_bf16(o_acc[8 * g + 4] * inv_l, o_acc[8 * g + 5] * inv_l);
                q.w = pack_bf16(o_acc[8 * g + 6] * inv_l, o_acc[8 * g + 7] * inv_l);
                *reinterpret_cast<uint4*>(o + 8 * g) = q;
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<1>(tmem_base, AT_TMEM_COLS);
}

}  // namespace kx

namespace kx {
int launch_attn_pp(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out, int batch,
                   int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out, cudaStream_t stream);   // attention_pp.cu
// KX_ATTN_IMPL=0 selects the first-generation kernel of this file (kept for A/B measurements).
static int attn_impl() {
    static int impl = -1;
    if (impl < 0) {
        const char* e = getenv("KX_ATTN_IMPL");
        impl = (e && e[0] == '0') ? 0 : 1;
    }
    return impl;
}
}  // namespace kx

using namespace kx;

static int attn_fwd_impl(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                         int batch, int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out,
                         cudaStream_t stream);

extern "C" int kx_attn_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                           int batch, int heads, int seq_len, int causal, float scale, float* stats_out,
                           cudaStream_t stream) {
    return attn_fwd_impl(q, k, v, ld_qkv, out, ld_out, batch, heads, seq_len, causal, scale, stats_out, nullptr, stream);
}

// Training forward: also writes the row log-sum-exp (log2 units, [heads][batch][ceil(T/128)*128] fp32) for kx_attn_bwd.
extern "C" int kx_attn_fwd_lse(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                               int batch, int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out,
                               cudaStream_t stream) {
    if (!lse_out || (reinterpret_cast<uintptr_t>(lse_out) & 15)) { set_error("kx_attn_fwd_lse: lse_out must be a 16-byte aligned buffer"); return KX_ERR_ARG; }
    return attn_fwd_impl(q, k, v, ld_qkv, out, ld_out, batch, heads, seq_len, causal, scale, stats_out, lse_out, stream);
}

static int attn_fwd_impl(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                         int batch, int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out,
                         cudaStream_t stream) {
    if (!q || !k || !v || !out) { set_error("kx_attn_fwd: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || heads <= 0 || seq_len <= 0) { set_error("kx_attn_fwd: bad shape"); return KX_ERR_ARG; }
    if ((ld_qkv % 8) || (ld_out % 8) || ((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((uintptr_t)out & 15)) {
        set_error("kx_attn_fwd: q/k/v/out need 16-byte aligned base and row pitch");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    if (stats_out && (reinterpret_cast<uintptr_t>(stats_out) & 7)) { set_error("kx_attn_fwd: stats_out must be 8-byte aligned"); return KX_ERR_ARG; }
    if (attn_impl() == 1) return launch_attn_pp(q, k, v, ld_qkv, out, ld_out, batch, heads, seq_len, causal, scale, stats_out, lse_out, stream);
    if (stats_out || lse_out) { set_error("kx_attn_fwd: stats_out is only produced by the default kernel (unset KX_ATTN_IMPL)"); return KX_ERR_ARG; }
    const unsigned long long rows = (unsigned long long)batch * seq_len;
    CUtensorMap tq, tk, tv;
    if (!make_tmap_bf16_2d(&tq, q, (uint64_t)heads * AT_D, rows, ld_qkv * 2, AT_D, AT_BM)) return KX_ERR_TMAP;
    if (!make_tmap_bf16_2d(&tk, k, (uint64_t)heads * AT_D, rows, ld_qkv * 2, AT_D, AT_BN)) return KX_ERR_TMAP;
    if (!make_tmap_bf16_2d(&tv, v, (uint64_t)heads * AT_D, rows, ld_qkv * 2, AT_D, AT_BN)) return KX_ERR_TMAP;
    AttnParams p;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ld_out = ld_out;
    p.seq_len = seq_len;
    p.heads = heads;
    p.num_q_tiles = (seq_len + AT_BM - 1) / AT_BM;
    p.scale_log2 = scale * 1.4426950408889634f;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
        cudaError_t e2 = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("kx_attn_fwd: cudaFuncSetAttribute failed"); return KX_ERR_LAUNCH; }
        attr_set = true;
    }
    dim3 grid(p.num_q_tiles, heads, batch);
    if (causal) attn_fwd_kernel<true><<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tq, tk, tv, p);
    else attn_fwd_kernel<false><<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tq, tk, tv, p);
    return check_launch("kx_attn_fwd");
}
