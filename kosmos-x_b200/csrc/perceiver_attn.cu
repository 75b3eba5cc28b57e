// kx_perceiver_xattn_fwd: the cross-attention core of flamingo_pytorch's PerceiverAttention
// (reached from /root/reference/kosmosx/model.py:231): 64 latent queries attend over
// [257 media tokens | 64 latents] per (batch, head), head_dim 64.  The forward runs on the tensor-core flash kernel of
// attention_pp.cu (kv_len != seq_len form: one item per (batch, head), three 128-key blocks, the unused query rows of the
// 128-row tiles masked at the store); the backward below is a CUDA-core kernel (2.6 MFLOP per (batch, head)).
#include "kx_internal.h"
#include "ptx.cuh"

namespace kx {
int launch_attn_pp(const void* q, long long ld_q, const void* k, const void* v, long long ld_kv, void* out, long long ld_out,
                   int batch, int heads, int seq_len, int kv_len, int causal, float scale, float* stats_out, float* lse_out,
                   cudaStream_t stream, float inv_keep, const uint32_t* row_mask);   // attention_pp.cu

constexpr int PX_KV_BLOCK = 128;
constexpr int PX_THREADS = 256;          // 64 query rows x 4 threads
}  // namespace kx

using namespace kx;

extern "C" int kx_perceiver_xattn_fwd(const void* q, long long ld_q, const void* kv, long long ld_kv, int v_col_off,
                                      void* out, long long ld_out, int batch, int heads, int n_q, int n_kv, float scale,
                                      cudaStream_t stream) {
    if (!q || !kv || !out) { set_error("kx_perceiver_xattn_fwd: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || heads <= 0 || n_q <= 0 || n_kv <= 0 || (ld_q % 8) || (ld_kv % 8) || (ld_out % 8) || (v_col_off % 8) ||
        ((uintptr_t)q & 15) || ((uintptr_t)kv & 15) || ((uintptr_t)out & 15)) {
        set_error("kx_perceiver_xattn_fwd: bad shape or alignment");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(kv);
    return launch_attn_pp(q, ld_q, k, k + v_col_off, ld_kv, out, ld_out, batch, heads, n_q, n_kv, 0, scale, nullptr, nullptr, stream,
                          1.0f, nullptr);
}

// ----------------------------------------------------------------------------- backward
// kx_perceiver_xattn_bwd: dq, dk, dv of the cross-attention above (training step; the resampler is trainable in the
// reference, model.py:196-203).  One CTA per (batch, head), n_q <= 64 queries.  Pass 1 recomputes the row maxima /
// sums; pass 2 walks 128-key blocks: with thread = (query row, 16-dim part) it rebuilds P and dS = P * (dP - delta)
// into shared memory and accumulates dq, then with thread = (key row, 32-dim half) it reduces dk = dS^T q and
// dv = P^T dO over the queries.  No atomics: every output element has one owner, results are bit-reproducible.
namespace kx {

constexpr int PXB_LD = PX_KV_BLOCK + 1;                 // padded row pitch of the P / dS blocks (fp32)
constexpr int PXB_SMEM = 2 * PX_KV_BLOCK * 64 * 2 + 2 * 64 * 64 * 2 + 2 * 64 * PXB_LD * 4;

__global__ void __launch_bounds__(PX_THREADS)
perceiver_xattn_bwd_kernel(const __nv_bfloat16* __restrict__ q, long long ld_q, const __nv_bfloat16* __restrict__ kv,
                           long long ld_kv, int v_col_off, const __nv_bfloat16* __restrict__ out, long long ld_out,
                           const __nv_bfloat16* __restrict__ d_out, long long ld_dout, __nv_bfloat16* __restrict__ dq,
                           long long ld_dq, __nv_bfloat16* __restrict__ dkv, long long ld_dkv, int n_q, int n_kv, float scale) {
    extern __shared__ __align__(16) uint8_t px_smem[];
    __nv_bfloat16 (*sk)[64] = reinterpret_cast<__nv_bfloat16 (*)[64]>(px_smem);
    __nv_bfloat16 (*sv)[64] = sk + PX_KV_BLOCK;
    __nv_bfloat16 (*sq)[64] = sv + PX_KV_BLOCK;         // [64][64] q (unscaled)
    __nv_bfloat16 (*sdo)[64] = sq + 64;                 // [64][64] dO
    float* sp = reinterpret_cast<float*>(sdo + 64);     // [64][PXB_LD] P block
    float* sds = sp + 64 * PXB_LD;                      // [64][PXB_LD] dS block
    const int head = blockIdx.x, b = blockIdx.y;
    const int qi = threadIdx.x >> 2, part = threadIdx.x & 3;
    const bool q_ok = qi < n_q;

    // stage q and dO (zero rows beyond n_q), keep this thread's 16 dims in registers
    for (int idx = threadIdx.x; idx < 64 * 8; idx += PX_THREADS) {
        const int row = idx >> 3, c = idx & 7;
        uint4 a = make_uint4(0, 0, 0, 0), d = make_uint4(0, 0, 0, 0);
        if (row < n_q) {
            a = *reinterpret_cast<const uint4*>(q + (static_cast<long long>(b) * n_q + row) * ld_q + head * 64 + c * 8);
            d = *reinterpret_cast<const uint4*>(d_out + (static_cast<long long>(b) * n_q + row) * ld_dout + head * 64 + c * 8);
        }
        *reinterpret_cast<uint4*>(&sq[row][c * 8]) = a;
        *reinterpret_cast<uint4*>(&sdo[row][c * 8]) = d;
    }
    __syncthreads();
    float qr[16], dor[16], delta = 0.f;
    {
        const __nv_bfloat16* op = out + (static_cast<long long>(b) * n_q + (q_ok ? qi : 0)) * ld_out + head * 64 + part * 16;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            qr[u] = __bfloat162float(sq[qi][part * 16 + u]) * scale;
            dor[u] = __bfloat162float(sdo[qi][part * 16 + u]);
            delta = fmaf(dor[u], q_ok ? __bfloat162float(op[u]) : 0.f, delta);
        }
        delta += __shfl_xor_sync(0xffffffffu, delta, 1);
        delta += __shfl_xor_sync(0xffffffffu, delta, 2);
    }
    auto stage_kv = [&](int kv0, int nk, bool with_v) {
        for (int idx = threadIdx.x; idx < PX_KV_BLOCK * 16; idx += PX_THREADS) {
            const int row = idx >> 4, c = idx & 15;
            if (c >= 8 && !with_v) continue;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (row < nk) {
                const __nv_bfloat16* src = kv + (static_cast<long long>(b) * n_kv + kv0 + row) * ld_kv + head * 64;
                val = (c < 8) ? *reinterpret_cast<const uint4*>(src + c * 8) : *reinterpret_cast<const uint4*>(src + v_col_off + (c - 8) * 8);
            }
            if (c < 8) *reinterpret_cast<uint4*>(&sk[row][c * 8]) = val;
            else *reinterpret_cast<uint4*>(&sv[row][(c - 8) * 8]) = val;
        }
    };
    auto dot16 = [&](const float (&a)[16], const __nv_bfloat16* row) {
        const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(row);
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float2 f = __bfloat1622float2(r2[u]);
            s = fmaf(a[2 * u], f.x, s);
            s = fmaf(a[2 * u + 1], f.y, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        return s;
    };
    // ---- pass 1: row max and sum of exp
    float m = -INFINITY, l = 0.f;
    for (int kv0 = 0; kv0 < n_kv; kv0 += PX_KV_BLOCK) {
        const int nk = min(PX_KV_BLOCK, n_kv - kv0);
        __syncthreads();
        stage_kv(kv0, nk, false);
        __syncthreads();
        for (int j = 0; j < nk; ++j) {
            const float s = dot16(qr, &sk[j][part * 16]);
            const float m_new = fmaxf(m, s);
            l = l * __expf(m - m_new) + __expf(s - m_new);
            m = m_new;
        }
    }
    const float inv_l = 1.0f / l;
    // ---- pass 2
    float dqr[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) dqr[u] = 0.f;
    const int kj = threadIdx.x >> 1, half = threadIdx.x & 1;       // phase B mapping: key row, 32-dim half
    for (int kv0 = 0; kv0 < n_kv; kv0 += PX_KV_BLOCK) {
        const int nk = min(PX_KV_BLOCK, n_kv - kv0);
        __syncthreads();
        stage_kv(kv0, nk, true);
        __syncthreads();
        // phase A: P, dS for this block; dq += dS . K
        for (int j = 0; j < nk; ++j) {
            const float s = dot16(qr, &sk[j][part * 16]);
            const float dp = dot16(dor, &sv[j][part * 16]);
            const float p = q_ok ? __expf(s - m) * inv_l : 0.f;
            const float ds = p * (dp - delta);
            if (part == 0) { sp[qi * PXB_LD + j] = p; sds[qi * PXB_LD + j] = ds; }
            const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(&sk[j][part * 16]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float2 f = __bfloat1622float2(kr[u]);
                dqr[2 * u] = fmaf(ds, f.x, dqr[2 * u]);
                dqr[2 * u + 1] = fmaf(ds, f.y, dqr[2 * u + 1]);
            }
        }
        __syncthreads();
        // phase B: dk_j = scale * sum_i dS_ij q_i ; dv_j = sum_i P_ij dO_i
        if (kj < nk) {
            float dk[32], dv[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) { dk[u] = 0.f; dv[u] = 0.f; }
            for (int i = 0; i < 64; ++i) {
                const float ds = sds[i * PXB_LD + kj], p = sp[i * PXB_LD + kj];
                const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&sq[i][half * 32]);
                const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&sdo[i][half * 32]);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const float2 fq = __bfloat1622float2(q2[u]), fd = __bfloat1622float2(d2[u]);
                    dk[2 * u] = fmaf(ds, fq.x, dk[2 * u]); dk[2 * u + 1] = fmaf(ds, fq.y, dk[2 * u + 1]);
                    dv[2 * u] = fmaf(p, fd.x, dv[2 * u]); dv[2 * u + 1] = fmaf(p, fd.y, dv[2 * u + 1]);
                }
            }
            __nv_bfloat16* dst = dkv + (static_cast<long long>(b) * n_kv + kv0 + kj) * ld_dkv + head * 64 + half * 32;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint4 a, c;
                a.x = pack_bf16(dk[8 * u] * scale, dk[8 * u + 1] * scale); a.y = pack_bf16(dk[8 * u + 2] * scale, dk[8 * u + 3] * scale);
                a.z = pack_bf16(dk[8 * u + 4] * scale, dk[8 * u + 5] * scale); a.w = pack_bf16(dk[8 * u + 6] * scale, dk[8 * u + 7] * scale);
                c.x = pack_bf16(dv[8 * u], dv[8 * u + 1]); c.y = pack_bf16(dv[8 * u + 2], dv[8 * u + 3]);
                c.z = pack_bf16(dv[8 * u + 4], dv[8 * u + 5]); c.w = pack_bf16(dv[8 * u + 6], dv[8 * u + 7]);
                *reinterpret_cast<uint4*>(dst + 8 * u) = a;
                *reinterpret_cast<uint4*>(dst + v_col_off + 8 * u) = c;
            }
        }
    }
    if (q_ok) {      // dq = scale * dS . K (qr carried the scale for the scores only)
        __nv_bfloat16* dst = dq + (static_cast<long long>(b) * n_q + qi) * ld_dq + head * 64 + part * 16;
        uint4 a, c;
        a.x = pack_bf16(dqr[0] * scale, dqr[1] * scale); a.y = pack_bf16(dqr[2] * scale, dqr[3] * scale);
        a.z = pack_bf16(dqr[4] * scale, dqr[5] * scale); a.w = pack_bf16(dqr[6] * scale, dqr[7] * scale);
        c.x = pack_bf16(dqr[8] * scale, dqr[9] * scale); c.y = pack_bf16(dqr[10] * scale, dqr[11] * scale);
        c.z = pack_bf16(dqr[12] * scale, dqr[13] * scale); c.w = pack_bf16(dqr[14] * scale, dqr[15] * scale);
        *reinterpret_cast<uint4*>(dst) = a;
        *reinterpret_cast<uint4*>(dst + 8) = c;
    }
}

}  // namespace kx

extern "C" int kx_perceiver_xattn_bwd(const void* q, long long ld_q, const void* kv, long long ld_kv, int v_col_off,
                                      const void* out, long long ld_out, const void* d_out, long long ld_dout, void* dq,
                                      long long ld_dq, void* dkv, long long ld_dkv, int batch, int heads, int n_q, int n_kv,
                                      float scale, cudaStream_t stream) {
    if (!q || !kv || !out || !d_out || !dq || !dkv) { set_error("kx_perceiver_xattn_bwd: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || heads <= 0 || n_q <= 0 || n_q > 64 || n_kv <= 0 || (ld_q % 8) || (ld_kv % 8) || (ld_out % 8) || (ld_dout % 8) ||
        (ld_dq % 8) || (ld_dkv % 8) || (v_col_off % 8) || ((uintptr_t)q & 15) || ((uintptr_t)kv & 15) || ((uintptr_t)out & 15) ||
        ((uintptr_t)d_out & 15) || ((uintptr_t)dq & 15) || ((uintptr_t)dkv & 15)) {
        set_error("kx_perceiver_xattn_bwd: bad shape or alignment (n_q <= 64, 16-byte aligned rows)");
        return KX_ERR_ARG;
    }
    if (kx::device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kx::perceiver_xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kx::PXB_SMEM);
        if (e != cudaSuccess) { kx::set_error("kx_perceiver_xattn_bwd: cudaFuncSetAttribute failed"); return KX_ERR_LAUNCH; }
        attr_set = true;
    }
    kx::perceiver_xattn_bwd_kernel<<<dim3(heads, batch), kx::PX_THREADS, kx::PXB_SMEM, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(q), ld_q, reinterpret_cast<const __nv_bfloat16*>(kv), ld_kv, v_col_off,
        reinterpret_cast<const __nv_bfloat16*>(out), ld_out, reinterpret_cast<const __nv_bfloat16*>(d_out), ld_dout,
        reinterpret_cast<__nv_bfloat16*>(dq), ld_dq, reinterpret_cast<__nv_bfloat16*>(dkv), ld_dkv, n_q, n_kv, scale);
    return kx::check_launch("kx_perceiver_xattn_bwd");
}
