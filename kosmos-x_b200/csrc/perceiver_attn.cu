// kx_perceiver_xattn_fwd: the cross-attention core of flamingo_pytorch's PerceiverAttention
// (reached from /root/reference/kosmosx/model.py:231).
//
// 64 latent queries attend over [257 media tokens ‖ 64 latents] per (batch, head), head_dim 64:
// 2.6 MFLOP per CTA — far below one tensor-core tile's worth of work, so this is a plain fp32
// CUDA-core kernel: K/V blocks staged in shared memory (coalesced 16-byte loads), four threads
// per query row each owning 16 of the 64 dims, online softmax in registers.
#include "kx_internal.h"
#include "ptx.cuh"

namespace kx {

constexpr int PX_KV_BLOCK = 128;
constexpr int PX_THREADS = 256;          // 64 query rows x 4 threads

__global__ void __launch_bounds__(PX_THREADS)
perceiver_xattn_kernel(const __nv_bfloat16* __restrict__ q, long long ld_q, const __nv_bfloat16* __restrict__ kv,
                       long long ld_kv, int v_col_off, __nv_bfloat16* __restrict__ out, long long ld_out, int n_q,
                       int n_kv, float scale) {
    __shared__ __align__(16) __nv_bfloat16 sk[PX_KV_BLOCK][64];
    __shared__ __align__(16) __nv_bfloat16 sv[PX_KV_BLOCK][64];
    const int head = blockIdx.x, b = blockIdx.y, qblk = blockIdx.z;
    const int qi = qblk * 64 + (threadIdx.x >> 2);      // query row
    const int part = threadIdx.x & 3;                   // which 16 dims
    const bool q_ok = qi < n_q;

    float qr[16];
    {
        const __nv_bfloat16* qp = q + (static_cast<long long>(b) * n_q + (q_ok ? qi : 0)) * ld_q + head * 64 + part * 16;
        const uint4 a = *reinterpret_cast<const uint4*>(qp);
        const uint4 c = *reinterpret_cast<const uint4*>(qp + 8);
        const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 f0 = __bfloat1622float2(h0[u]), f1 = __bfloat1622float2(h1[u]);
            qr[2 * u] = f0.x * scale; qr[2 * u + 1] = f0.y * scale;
            qr[8 + 2 * u] = f1.x * scale; qr[8 + 2 * u + 1] = f1.y * scale;
        }
    }
    float o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;

    for (int kv0 = 0; kv0 < n_kv; kv0 += PX_KV_BLOCK) {
        const int nk = min(PX_KV_BLOCK, n_kv - kv0);
        __syncthreads();
        // stage K and V rows: 8 x 16-byte chunks per row each
        for (int idx = threadIdx.x; idx < nk * 16; idx += PX_THREADS) {
            const int row = idx >> 4, c = idx & 15;
            const __nv_bfloat16* src = kv + (static_cast<long long>(b) * n_kv + kv0 + row) * ld_kv + head * 64;
            if (c < 8) *reinterpret_cast<uint4*>(&sk[row][c * 8]) = *reinterpret_cast<const uint4*>(src + c * 8);
            else *reinterpret_cast<uint4*>(&sv[row][(c - 8) * 8]) = *reinterpret_cast<const uint4*>(src + v_col_off + (c - 8) * 8);
        }
        __syncthreads();
        for (int j = 0; j < nk; ++j) {
            const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(&sk[j][part * 16]);
            float s = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float2 f = __bfloat1622float2(kr[u]);
                s = fmaf(qr[2 * u], f.x, s);
                s = fmaf(qr[2 * u + 1], f.y, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const float m_new = fmaxf(m, s);
            const float alpha = __expf(m - m_new);
            const float p = __expf(s - m_new);
            l = l * alpha + p;
            const __nv_bfloat162* vr = reinterpret_cast<const __nv_bfloat162*>(&sv[j][part * 16]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float2 f = __bfloat1622float2(vr[u]);
                o[2 * u] = fmaf(p, f.x, o[2 * u] * alpha);
                o[2 * u + 1] = fmaf(p, f.y, o[2 * u + 1] * alpha);
            }
            m = m_new;
        }
    }
    if (q_ok) {
        const float inv = 1.0f / l;
        __nv_bfloat16* op = out + (static_cast<long long>(b) * n_q + qi) * ld_out + head * 64 + part * 16;
        uint4 a, c;
        a.x = pack_bf16(o[0] * inv, o[1] * inv); a.y = pack_bf16(o[2] * inv, o[3] * inv);
        a.z = pack_bf16(o[4] * inv, o[5] * inv); a.w = pack_bf16(o[6] * inv, o[7] * inv);
        c.x = pack_bf16(o[8] * inv, o[9] * inv); c.y = pack_bf16(o[10] * inv, o[11] * inv);
        c.z = pack_bf16(o[12] * inv, o[13] * inv); c.w = pack_bf16(o[14] * inv, o[15] * inv);
        *reinterpret_cast<uint4*>(op) = a;
        *reinterpret_cast<uint4*>(op + 8) = c;
    }
}

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
    dim3 grid(heads, batch, (n_q + 63) / 64);
    perceiver_xattn_kernel<<<grid, PX_THREADS, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(q), ld_q, reinterpret_cast<const __nv_bfloat16*>(kv), ld_kv, v_col_off,
        reinterpret_cast<__nv_bfloat16*>(out), ld_out, n_q, n_kv, scale);
    return check_launch("kx_perceiver_xattn_fwd");
}
