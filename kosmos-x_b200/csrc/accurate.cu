// Verification-precision kernels of the Kosmos-X path ("bf16x3" mode of kosmosx.Kosmos).
//
// The throughput path rounds every tensor-core operand to bf16 (one ulp at 1.0 = 7.8e-3), so its logits sit ~4e-2 from
// the fp32 reference.  To SHOW that the kernels compute the reference's function to the tolerance BASELINE.json states
// (logits max-abs-diff <= 1e-3 against /root/reference/kosmosx/model.py:208-253 run in fp32) the same tcgen05 GEMM is fed
// split operands: an fp32 value x is written as hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits), and
//     A.W^T ~= Ah.Wh^T + Ah.Wl^T + Al.Wh^T
// is ONE kx_gemm_bf16 launch over K' = 3K: activations are laid out [hi | hi | lo], weights [hi | lo | hi], so the three
// products accumulate in the same fp32 TMEM tile (the dropped Al.Wl term is 2^-18 relative).  Everything between the GEMMs
// stays fp32: kx_layernorm_fwd (fp32 in / fp32 out), the fp32 GEMM epilogues (bias, GELU, residual), and the three small
// kernels below — attention in plain fp32 FMAs (the tensor-core flash kernels take bf16 P and bf16 q/k/v), the xPos
// rotation on fp32 q|k, and the patch im2col in fp32.  This mode is for parity, not speed.
#include "kx_internal.h"
#include "ptx.cuh"

namespace kx {

// dst[r, s*n_pad + k]: weights = 0 -> (hi, hi, lo) for s = 0, 1, 2; weights = 1 -> (hi, lo, hi).  k in [n, n_pad) is zero.
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ src, long long ld_src, int rows, int n, int n_pad,
                    __nv_bfloat16* __restrict__ dst, long long ld_dst, int weights) {
    const long long total = static_cast<long long>(rows) * n_pad;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / n_pad;
        const int k = static_cast<int>(i - r * n_pad);
        const float x = k < n ? src[r * ld_src + k] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        const __nv_bfloat16 lo = __float2bfloat16_rn(__fsub_rn(x, __bfloat162float(hi)));
        __nv_bfloat16* o = dst + r * ld_dst + k;
        o[0] = hi;
        o[n_pad] = weights ? lo : hi;
        o[2 * static_cast<long long>(n_pad)] = weights ? hi : lo;
    }
}

// softmax(q.k^T * scale [+ causal mask]) . v in fp32, head_dim 64.  q rows [batch*n_q, ld_q], k / v rows [batch*n_kv, ld_kv],
// head h at columns h*64..h*64+63 of the given base pointers.  One CTA = 32 query rows of one (batch, head): a warp owns 4
// rows, lane = key inside a 32-key chunk staged in shared memory (k padded to 65 floats per row: conflict-free), online
// softmax with expf, P.V accumulated with lane = output column (d = lane, lane + 32).  Replaces, at verification
// precision, the same reference code as kx_attn_fwd and kx_perceiver_xattn_fwd.
constexpr int AF_ROWS = 32, AF_KEYS = 32;
__global__ void __launch_bounds__(256)
attn_f32_kernel(const float* __restrict__ q, long long ld_q, const float* __restrict__ k, const float* __restrict__ v,
                long long ld_kv, float* __restrict__ out, long long ld_out, int n_q, int n_kv, int causal, float scale) {
    __shared__ float sq[AF_ROWS][64];
    __shared__ float sk[AF_KEYS][65];
    __shared__ float sv[AF_KEYS][64];
    const int b = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * AF_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* qb = q + (static_cast<long long>(b) * n_q) * ld_q + h * 64;
    const float* kb = k + (static_cast<long long>(b) * n_kv) * ld_kv + h * 64;
    const float* vb = v + (static_cast<long long>(b) * n_kv) * ld_kv + h * 64;
    for (int i = threadIdx.x; i < AF_ROWS * 64; i += 256) {
        const int r = i >> 6, d = i & 63;
        sq[r][d] = (r0 + r < n_q) ? qb[static_cast<long long>(r0 + r) * ld_q + d] : 0.f;
    }
    float m[4], l[4], o0[4], o1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { m[u] = -INFINITY; l[u] = 0.f; o0[u] = 0.f; o1[u] = 0.f; }
    const int last_row = min(r0 + AF_ROWS, n_q) - 1;
    const int kv_end = causal ? min(n_kv, last_row + 1) : n_kv;      // causal: key j <= query row
    for (int j0 = 0; j0 < kv_end; j0 += AF_KEYS) {
        __syncthreads();
        for (int i = threadIdx.x; i < AF_KEYS * 64; i += 256) {
            const int j = i >> 6, d = i & 63;
            const bool in = j0 + j < n_kv;
            sk[j][d] = in ? kb[static_cast<long long>(j0 + j) * ld_kv + d] : 0.f;
            sv[j][d] = in ? vb[static_cast<long long>(j0 + j) * ld_kv + d] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rl = warp * 4 + u, row = r0 + rl;
            if (row >= n_q) continue;                                 // warp-uniform
            const int key = j0 + lane;
            float s = 0.f;
#pragma unroll 16
            for (int d = 0; d < 64; ++d) s = fmaf(sq[rl][d], sk[lane][d], s);
            s *= scale;
            const bool valid = key < n_kv && (!causal || key <= row);
            s = valid ? s : -INFINITY;
            float cm = s;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, off));
            const float m_new = fmaxf(m[u], cm);
            if (m_new == -INFINITY) continue;                         // whole chunk masked for this row (warp-uniform)
            const float p = valid ? expf(s - m_new) : 0.f;
            const float corr = expf(m[u] - m_new);                    // m[u] = -inf -> 0
            float ps = p;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
            l[u] = l[u] * corr + ps;
            float a0 = o0[u] * corr, a1 = o1[u] * corr;
#pragma unroll 8
            for (int j = 0; j < AF_KEYS; ++j) {
                const float pj = __shfl_sync(0xffffffffu, p, j);
                a0 = fmaf(pj, sv[j][lane], a0);
                a1 = fmaf(pj, sv[j][lane + 32], a1);
            }
            o0[u] = a0; o1[u] = a1; m[u] = m_new;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int row = r0 + warp * 4 + u;
        if (row >= n_q) continue;
        float* po = out + (static_cast<long long>(b) * n_q + row) * ld_out + h * 64;
        const float inv = 1.0f / l[u];
        po[lane] = o0[u] * inv;
        po[lane + 32] = o1[u] * inv;
    }
}

// xPos rotation (same tables and pairing as the KX_EPI_QKV_XPOS epilogue) in place on the q and k column
// blocks of an fp32 [rows, ld] q|k|v matrix: pair j of a head = columns (2j, 2j+1); out0 = x0*c - x1*s, out1 = x1*c + x0*s.
__global__ void __launch_bounds__(256)
xpos_apply_f32_kernel(float* __restrict__ qkv, long long ld, int rows, int d_model, int seq_len, const float* __restrict__ q_cos,
                      const float* __restrict__ q_sin, const float* __restrict__ k_cos, const float* __restrict__ k_sin) {
    const int pairs = d_model;                       // (2 * d_model) / 2 pairs per row over the q and k blocks
    const long long total = static_cast<long long>(rows) * pairs;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / pairs;
        const int pc = static_cast<int>(i - r * pairs);
        const int col = 2 * pc;                      // column inside [0, 2*d_model)
        const bool is_k = col >= d_model;
        const int j = (col & 63) >> 1;
        const int t = static_cast<int>(r % seq_len);
        const float c = (is_k ? k_cos : q_cos)[t * 32 + j];
        const float s = (is_k ? k_sin : q_sin)[t * 32 + j];
        float2* p = reinterpret_cast<float2*>(qkv + r * ld + col);
        const float2 x = *p;
        *p = make_float2(x.x * c - x.y * s, x.y * c + x.x * s);
    }
}

// fp32 twin of kx_im2col_patches ([HF] modeling_clip.py:202-218): patches[slot*P + p, k] = pixels[n, c, py*patch + dy,
// px*patch + dx], k = c*patch*patch + dy*patch + dx (zero padded to k_pad), and the CLS rows x[slot, 0, :] = class_embedding
// + pos[0].  slot = i*(N/media) + s for image i of sequence s (media-major), n = s*media + i.
__global__ void __launch_bounds__(256)
im2col_f32_kernel(const float* __restrict__ pixels, int batch, int media, int image, int patch, float* __restrict__ patches,
                  int k_pad, const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x, int dim) {
    const int g = image / patch, P = g * g, kk = 3 * patch * patch;
    const long long n_patch = static_cast<long long>(batch) * P * k_pad;
    const long long total = n_patch + static_cast<long long>(batch) * dim;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const int seqs = batch / media;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (i < n_patch) {
            const int k = static_cast<int>(i % k_pad);
            const long long rp = i / k_pad;
            const int p = static_cast<int>(rp % P);
            const int slot = static_cast<int>(rp / P);
            const int n = (slot % seqs) * media + slot / seqs;
            float val = 0.f;
            if (k < kk) {
                const int c = k / (patch * patch), rem = k - c * patch * patch;
                const int dy = rem / patch, dx = rem - dy * patch;
                const int py = p / g, px = p - py * g;
                val = pixels[((static_cast<long long>(n) * 3 + c) * image + py * patch + dy) * image + px * patch + dx];
            }
            patches[i] = val;
        } else {
            const long long e = i - n_patch;
            const int slot = static_cast<int>(e / dim), d = static_cast<int>(e - static_cast<long long>(slot) * dim);
            x[static_cast<long long>(slot) * (P + 1) * dim + d] = cls[d] + pos[d];
        }
    }
}

static int grid_for(long long total) {
    const long long blocks = (total + 255) / 256;
    return static_cast<int>(std::min<long long>(blocks, 148ll * 16));
}

}  // namespace kx

using namespace kx;

extern "C" int kx_split_bf16x3(const float* src, long long ld_src, int rows, int n, int n_pad, void* dst_bf16, long long ld_dst,
                               int weights, cudaStream_t stream) {
    if (!src || !dst_bf16) { set_error("kx_split_bf16x3: null pointer"); return KX_ERR_ARG; }
    if (rows <= 0 || n <= 0 || n_pad < n || (n_pad % 8) || ld_dst < 3ll * n_pad || (ld_dst % 8) ||
        (reinterpret_cast<uintptr_t>(dst_bf16) & 15)) {
        set_error("kx_split_bf16x3: bad shape (rows=%d n=%d n_pad=%d ld_dst=%lld); n_pad %% 8 == 0, ld_dst >= 3*n_pad", rows, n, n_pad, ld_dst);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    split_bf16x3_kernel<<<grid_for(static_cast<long long>(rows) * n_pad), 256, 0, stream>>>(
        src, ld_src, rows, n, n_pad, reinterpret_cast<__nv_bfloat16*>(dst_bf16), ld_dst, weights ? 1 : 0);
    return check_launch("kx_split_bf16x3");
}

extern "C" int kx_attn_f32(const float* q, long long ld_q, const float* k, const float* v, long long ld_kv, float* out,
                           long long ld_out, int batch, int heads, int n_q, int n_kv, int causal, float scale,
                           cudaStream_t stream) {
    if (!q || !k || !v || !out) { set_error("kx_attn_f32: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || heads <= 0 || n_q <= 0 || n_kv <= 0 || batch > 65535 || heads > 65535 || (causal && n_q != n_kv)) {
        set_error("kx_attn_f32: bad shape (batch=%d heads=%d n_q=%d n_kv=%d causal=%d)", batch, heads, n_q, n_kv, causal);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    dim3 grid((n_q + AF_ROWS - 1) / AF_ROWS, heads, batch);
    attn_f32_kernel<<<grid, 256, 0, stream>>>(q, ld_q, k, v, ld_kv, out, ld_out, n_q, n_kv, causal ? 1 : 0, scale);
    return check_launch("kx_attn_f32");
}

extern "C" int kx_xpos_apply_f32(float* qkv, long long ld, int rows, int d_model, int seq_len, const float* q_cos,
                                 const float* q_sin, const float* k_cos, const float* k_sin, cudaStream_t stream) {
    if (!qkv || !q_cos || !q_sin || !k_cos || !k_sin) { set_error("kx_xpos_apply_f32: null pointer"); return KX_ERR_ARG; }
    if (rows <= 0 || d_model <= 0 || (d_model % 64) || seq_len <= 0 || (ld % 2) || ld < 2ll * d_model ||
        (reinterpret_cast<uintptr_t>(qkv) & 7)) {
        set_error("kx_xpos_apply_f32: bad shape (rows=%d d_model=%d seq_len=%d ld=%lld)", rows, d_model, seq_len, ld);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    xpos_apply_f32_kernel<<<grid_for(static_cast<long long>(rows) * d_model), 256, 0, stream>>>(qkv, ld, rows, d_model, seq_len, q_cos,
                                                                                                q_sin, k_cos, k_sin);
    return check_launch("kx_xpos_apply_f32");
}

extern "C" int kx_im2col_patches_f32(const float* pixels, int batch, int media, int image, int patch, float* patches, int k_pad,
                                     const float* class_embedding, const float* pos_table, float* x, int dim,
                                     cudaStream_t stream) {
    if (!pixels || !patches || !class_embedding || !pos_table || !x) { set_error("kx_im2col_patches_f32: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || media <= 0 || (batch % media) || image <= 0 || patch <= 0 || (image % patch) || k_pad < 3 * patch * patch || dim <= 0) {
        set_error("kx_im2col_patches_f32: bad shape (batch=%d media=%d image=%d patch=%d k_pad=%d)", batch, media, image, patch, k_pad);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    const int g = image / patch;
    const long long total = static_cast<long long>(batch) * g * g * k_pad + static_cast<long long>(batch) * dim;
    im2col_f32_kernel<<<grid_for(total), 256, 0, stream>>>(pixels, batch, media, image, patch, patches, k_pad, class_embedding,
                                                           pos_table, x, dim);
    return check_launch("kx_im2col_patches_f32");
}
