// kx_attn_fwd / kx_attn_fwd_lse: argument checks of the flash-attention forward entry points; the kernel is the persistent
// ping-pong tcgen05 kernel of attention_pp.cu (S / P / O in TMEM, FA4-style).  (The first-generation kernel that used to
// live here behind an environment switch was removed in round 2: one code path, no dispatch.)
#include "kx_internal.h"
#include "philox.cuh"

namespace kx {
int launch_attn_pp(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out, int batch,
                   int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out, cudaStream_t stream,
                   const DropSpec* drop, uint32_t* drop_mask);   // attention_pp.cu
}  // namespace kx

using namespace kx;

static int attn_fwd_impl(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                         int batch, int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out,
                         cudaStream_t stream, const DropSpec* drop = nullptr, uint32_t* drop_mask = nullptr) {
    if (!q || !k || !v || !out) { set_error("kx_attn_fwd: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || heads <= 0 || seq_len <= 0) { set_error("kx_attn_fwd: bad shape"); return KX_ERR_ARG; }
    if ((ld_qkv % 8) || (ld_out % 8) || ((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((uintptr_t)out & 15)) {
        set_error("kx_attn_fwd: q/k/v/out need 16-byte aligned base and row pitch");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    if (stats_out && (reinterpret_cast<uintptr_t>(stats_out) & 7)) { set_error("kx_attn_fwd: stats_out must be 8-byte aligned"); return KX_ERR_ARG; }
    return launch_attn_pp(q, k, v, ld_qkv, out, ld_out, batch, heads, seq_len, causal, scale, stats_out, lse_out, stream, drop, drop_mask);
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

// Training forward with attention dropout (see the header): the same kernel, DROP instantiation.
extern "C" int kx_attn_fwd_dropout(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                                   int batch, int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out,
                                   float drop_p, unsigned int drop_site, unsigned long long drop_seed, unsigned int* drop_mask,
                                   cudaStream_t stream) {
    if (!lse_out || (reinterpret_cast<uintptr_t>(lse_out) & 15)) { set_error("kx_attn_fwd_dropout: lse_out must be a 16-byte aligned buffer"); return KX_ERR_ARG; }
    if (!(drop_p > 0.f && drop_p < 1.f) || !drop_mask || (reinterpret_cast<uintptr_t>(drop_mask) & 15) || !causal) {
        set_error("kx_attn_fwd_dropout: needs 0 < p < 1, causal attention and a 16-byte aligned mask buffer");
        return KX_ERR_ARG;
    }
    const DropSpec d = make_drop_spec(drop_p, drop_site, drop_seed);
    return attn_fwd_impl(q, k, v, ld_qkv, out, ld_out, batch, heads, seq_len, causal, scale, stats_out, lse_out, stream, &d, drop_mask);
}
