// kx_attn_fwd / kx_attn_fwd_lse: argument checks of the flash-attention forward entry points; the kernel is the persistent
// ping-pong tcgen05 kernel of attention_pp.cu (S / P / O in TMEM, FA4-style).  (The first-generation kernel that used to
// live here behind an environment switch was removed in round 2: one code path, no dispatch.)
#include "kx_internal.h"
#include "philox.cuh"

namespace kx {
int launch_attn_pp(const void* q, long long ld_q, const void* k, const void* v, long long ld_kv, void* out, long long ld_out,
                   int batch, int heads, int seq_len, int kv_len, int causal, float scale, float* stats_out, float* lse_out,
                   cudaStream_t stream, float inv_keep, const uint32_t* row_mask);   // attention_pp.cu

// keep probability as the 12-bit fixed-point fraction the bit-sliced mask generator realises exactly
inline unsigned attn_keep_thr12(float drop_p) { return static_cast<unsigned>((1.0 - static_cast<double>(drop_p)) * 4096.0 + 0.5); }

// One CTA per 128 x 128 tile of one (batch, head): thread = query row.  Four 32-key words per row go out row-major (the
// forward kernel's thread = query row reads its uint4) and, transposed inside the warp, key-major (the backward kernel's
// thread = key reads one word per query quarter).
// The forward kernel masks P as bf16 PAIRS (keys 2j, 2j+1 share a register), so the keep bit of key k = 2j + e of a word is
// DEFINED to be bit (j % 8) + 8 e + 16 (j / 8) of the Philox word: inside each 16-key half the even keys' bits sit in the
// low byte and the odd keys' in the high byte, one shift puts both bits of a pair on the sign positions of two bytes and
// one PRMT (sign-replicate mode) expands them to the pair mask.  The bits are i.i.d., so naming them this way costs nothing:
// the row-major word is the raw word, and the transposed copy only stores its columns at the permuted key index.
__global__ void __launch_bounds__(128)
attn_dropout_mask_kernel(unsigned long long seed, uint32_t site, uint32_t thr12, int nb, int causal, uint4* __restrict__ row_mask,
                         uint32_t* __restrict__ key_mask) {
    const int tiles_per_bh = causal ? nb * (nb + 1) / 2 : nb * nb;
    const int bh = blockIdx.x / tiles_per_bh;
    int t = blockIdx.x - bh * tiles_per_bh;
    int qb, kb;
    if (causal) {                       // t-th (qb, kb <= qb) pair in row-major order of the lower triangle
        qb = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
        while (qb * (qb + 1) / 2 > t) --qb;
        while ((qb + 1) * (qb + 2) / 2 <= t) ++qb;
        kb = t - qb * (qb + 1) / 2;
    } else {
        qb = t / nb; kb = t - qb * nb;
    }
    const int r = threadIdx.x, lane = r & 31, g = r >> 5;
    const uint32_t q = static_cast<uint32_t>(qb * 128 + r);
    uint32_t w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) w[c] = drop_keep32_attn(seed, site, thr12, q, static_cast<uint32_t>(kb * 4 + c), static_cast<uint32_t>(bh));
    const long long tile = (static_cast<long long>(bh) * nb + qb) * nb + kb;
    row_mask[tile * 128 + r] = make_uint4(w[0], w[1], w[2], w[3]);
    // key-major copy: 32 x 32 bit-matrix transpose inside the warp (5 butterfly steps of one shuffle each) — lane b ends up
    // with bit i = (bit b of query 32g + i's word c) = keep(query 32g + i, key 32c + key_of_bit(b))
    uint32_t* dst = key_mask + tile * 512 + g * 128;
    const int key_of_bit = 2 * ((lane >> 4) * 8 + (lane & 7)) + ((lane >> 3) & 1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t x = w[c], m = 0x0000ffffu;
#pragma unroll
        for (int j = 16; j >= 1; j >>= 1) {
            const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
            x = (lane & j) ? ((x & ~m) | ((y & ~m) >> j)) : ((x & m) | ((y & m) << j));
            if (j > 1) m ^= m << (j >> 1);
        }
        dst[c * 32 + key_of_bit] = x;
    }
}
}  // namespace kx

using namespace kx;

static int attn_fwd_impl(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                         int batch, int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out,
                         cudaStream_t stream, float inv_keep = 1.0f, const uint32_t* row_mask = nullptr) {
    if (!q || !k || !v || !out) { set_error("kx_attn_fwd: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || heads <= 0 || seq_len <= 0) { set_error("kx_attn_fwd: bad shape"); return KX_ERR_ARG; }
    if ((ld_qkv % 8) || (ld_out % 8) || ((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((uintptr_t)out & 15)) {
        set_error("kx_attn_fwd: q/k/v/out need 16-byte aligned base and row pitch");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    if (stats_out && (reinterpret_cast<uintptr_t>(stats_out) & 7)) { set_error("kx_attn_fwd: stats_out must be 8-byte aligned"); return KX_ERR_ARG; }
    return launch_attn_pp(q, ld_qkv, k, v, ld_qkv, out, ld_out, batch, heads, seq_len, seq_len, causal, scale, stats_out, lse_out, stream,
                          inv_keep, row_mask);
}

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

extern "C" size_t kx_attn_dropout_mask_words(int batch, int heads, int seq_len) {
    if (batch <= 0 || heads <= 0 || seq_len <= 0) return 0;
    const size_t nb = (static_cast<size_t>(seq_len) + 127) / 128;
    return static_cast<size_t>(batch) * heads * nb * nb * 512;
}

// Keep bits of one attention-dropout site (see the header).  Both layouts are written: row_mask for kx_attn_fwd_dropout,
// key_mask for kx_attn_bwd_dropout; kx_attn_dropout_mask_words() words each.
extern "C" int kx_attn_dropout_masks(float drop_p, unsigned int drop_site, unsigned long long drop_seed, int batch, int heads,
                                     int seq_len, int causal, unsigned int* row_mask, unsigned int* key_mask, cudaStream_t stream) {
    if (!(drop_p > 0.f && drop_p < 1.f) || !row_mask || !key_mask || (reinterpret_cast<uintptr_t>(row_mask) & 15) ||
        (reinterpret_cast<uintptr_t>(key_mask) & 15) || batch <= 0 || heads <= 0 || seq_len <= 0) {
        set_error("kx_attn_dropout_masks: needs 0 < p < 1 and two 16-byte aligned buffers of kx_attn_dropout_mask_words() words");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    const int nb = (seq_len + 127) / 128;
    const long long tiles = static_cast<long long>(batch) * heads * (causal ? nb * (nb + 1) / 2 : nb * nb);
    if (tiles > 0x7fffffffLL) { set_error("kx_attn_dropout_masks: too many tiles"); return KX_ERR_ARG; }
    attn_dropout_mask_kernel<<<static_cast<unsigned>(tiles), 128, 0, stream>>>(drop_seed, drop_site, attn_keep_thr12(drop_p), nb,
                                                                               causal ? 1 : 0, reinterpret_cast<uint4*>(row_mask), key_mask);
    return check_launch("kx_attn_dropout_masks");
}

// Training forward with attention dropout (see the header): the same kernel, DROP instantiation, keep bits from row_mask.
extern "C" int kx_attn_fwd_dropout(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                                   int batch, int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out,
                                   float drop_p, const unsigned int* row_mask, cudaStream_t stream) {
    if (!lse_out || (reinterpret_cast<uintptr_t>(lse_out) & 15)) { set_error("kx_attn_fwd_dropout: lse_out must be a 16-byte aligned buffer"); return KX_ERR_ARG; }
    if (!(drop_p > 0.f && drop_p < 1.f) || !row_mask || (reinterpret_cast<uintptr_t>(row_mask) & 15) || !causal) {
        set_error("kx_attn_fwd_dropout: needs 0 < p < 1, causal attention and the 16-byte aligned row_mask of kx_attn_dropout_masks");
        return KX_ERR_ARG;
    }
    const float inv_keep = 4096.0f / static_cast<float>(attn_keep_thr12(drop_p));
    return attn_fwd_impl(q, k, v, ld_qkv, out, ld_out, batch, heads, seq_len, causal, scale, stats_out, lse_out, stream, inv_keep, row_mask);
}
