// Counter-based random bits for dropout (the reference trains with dropout = attention_dropout = 0.1,
// /root/reference/kosmosx/model.py:175-177): Philox4x32 (Salmon et al., SC'11) with 7 rounds (the paper's
// Crush-resistant minimum; torch uses 10).  One call -> 128 bits = eight 16-bit lots; an element is KEPT when its lot
// is below `thr` = round(keep * 65536), so the drop probability is exact to 2^-16 and forward and backward regenerate
// the same mask from (seed, site, coordinates) alone — nothing is stored for the element-wise sites.
#pragma once
#include <cstdint>

namespace kx {

struct DropSpec {
    unsigned long long seed;   // per training step
    unsigned int site;         // which dropout of the model (layer * 4 + kind); part of the counter
    unsigned int thr;          // keep if lot < thr; 0 = dropout off
    float inv_keep;            // 65536 / thr
};

inline DropSpec make_drop_spec(float p, unsigned int site, unsigned long long seed) {
    DropSpec d = {};
    if (p > 0.f && p < 1.f) {
        d.seed = seed;
        d.site = site;
        d.thr = static_cast<unsigned int>((1.0 - static_cast<double>(p)) * 65536.0 + 0.5);
        d.inv_keep = 65536.0f / static_cast<float>(d.thr);
    }
    return d;
}

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
        const uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// keep-mask (bit u = element u kept) of the 8 elements (row, 8*col8 .. 8*col8+7) of a 2-D dropout site
__device__ __forceinline__ uint32_t drop_keep8(const DropSpec& d, uint32_t row, uint32_t col8) {
    const uint4 r = philox4x32_7(row, col8, d.site, 0x6b78u, static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t m = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        m |= ((w[u] & 0xffffu) < d.thr ? 1u : 0u) << (2 * u);
        m |= ((w[u] >> 16) < d.thr ? 1u : 0u) << (2 * u + 1);
    }
    return m;
}

// Attention-probability site.  The keep bits of 32 consecutive keys of one query row come out of 12 uniform random
// words by bit-slicing: with thr12 = round(keep * 4096) = sum b_i 2^i, fold m = b_i ? (m | u_i) : (m & u_i) from i = 0 up —
// every bit of m is then Bernoulli(thr12 / 4096) — three Philox calls and 12 logic ops per 32 elements instead of 32
// sixteen-bit comparisons.  (q, key group, bh) address the words, so any kernel can regenerate any word.
__device__ __forceinline__ uint32_t drop_keep32_attn(unsigned long long seed, uint32_t site, uint32_t thr12, uint32_t q, uint32_t kgroup,
                                                    uint32_t bh) {
    uint32_t u[12];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint4 r = philox4x32_7(q, (kgroup << 2) | static_cast<uint32_t>(c), site, 0x61740000u + bh, static_cast<uint32_t>(seed),
                                     static_cast<uint32_t>(seed >> 32));
        u[4 * c] = r.x; u[4 * c + 1] = r.y; u[4 * c + 2] = r.z; u[4 * c + 3] = r.w;
    }
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < 12; ++i) m = ((thr12 >> i) & 1u) ? (m | u[i]) : (m & u[i]);
    return m;
}

}  // namespace kx
