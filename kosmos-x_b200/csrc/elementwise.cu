// HBM-bound glue kernels of the Kosmos-X path: LayerNorm (-> bf16 GEMM operand), token
// embedding + splice + positions, CLIP patch im2col, xPos tables, parameter staging.
// All are coalesced, 16-byte vectorised, one pass over their input.
#include "kx_internal.h"
#include "ptx.cuh"

namespace kx {

// ----------------------------------------------------------------------------- LayerNorm
// Row statistics exactly as torch.nn.LayerNorm: mean, then biased variance of (x - mean), eps
// inside the sqrt; fp32 throughout; the row lives in registers between the passes.
// EXACT: n == TPR * MAX_VEC * 8, so every load is unconditional and all of a thread's 16-byte loads are
// issued back to back before the first use (memory-level parallelism is what bounds this kernel).
template <bool IN_BF16, bool OUT_F32, int TPR, int MAX_VEC, bool EXACT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, long long ld_x, const float* __restrict__ pre_add_tab, int pre_add_group,
                 int pre_add_rows, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 void* __restrict__ out, long long ld_out, int rows, int n, int grp_rows, int grp_stride,
                 int grp_off) {
    constexpr int ROWS_PER_BLOCK = 256 / TPR;
    __shared__ float red[2][8];
    const int tr = threadIdx.x % TPR;
    const int row = blockIdx.x * ROWS_PER_BLOCK + threadIdx.x / TPR;
    const bool active = row < rows;          // inactive threads still take part in the reductions
    const int nvec = n >> 3;
    // pre_add row = (row / pre_add_group) % pre_add_rows  (perceiver media_pos_emb[:m])
    const float* pre_add = nullptr;
    if (pre_add_tab != nullptr && active)
        pre_add = pre_add_tab + static_cast<long long>(pre_add_group > 0 ? (row / pre_add_group) % pre_add_rows : 0) * n;

    float v[MAX_VEC][8];
    float sum = 0.f;
    if constexpr (EXACT) {
        const long long lrow = active ? row : (rows - 1);      // inactive threads re-read the last row, never store
        if constexpr (IN_BF16) {
            uint4 raw[MAX_VEC];
            const uint4* px = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + lrow * ld_x);
#pragma unroll
            for (int i = 0; i < MAX_VEC; ++i) raw[i] = px[tr + i * TPR];
#pragma unroll
            for (int i = 0; i < MAX_VEC; ++i) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[i]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 f = __bfloat1622float2(h[u]);
                    v[i][2 * u] = f.x; v[i][2 * u + 1] = f.y;
                }
            }
        } else {
            const float4* px = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + lrow * ld_x);
            float4 raw[MAX_VEC][2];
#pragma unroll
            for (int i = 0; i < MAX_VEC; ++i) {
                raw[i][0] = px[(tr + i * TPR) * 2];
                raw[i][1] = px[(tr + i * TPR) * 2 + 1];
            }
#pragma unroll
            for (int i = 0; i < MAX_VEC; ++i) {
                v[i][0] = raw[i][0].x; v[i][1] = raw[i][0].y; v[i][2] = raw[i][0].z; v[i][3] = raw[i][0].w;
                v[i][4] = raw[i][1].x; v[i][5] = raw[i][1].y; v[i][6] = raw[i][1].z; v[i][7] = raw[i][1].w;
            }
        }
        if (pre_add_tab != nullptr) {
            const float* pa = pre_add_tab + static_cast<long long>(pre_add_group > 0 ? (lrow / pre_add_group) % pre_add_rows : 0) * n;
#pragma unroll
            for (int i = 0; i < MAX_VEC; ++i) {
                const int vi = tr + i * TPR;
                const float4 a = __ldg(reinterpret_cast<const float4*>(pa + vi * 8));
                const float4 b = __ldg(reinterpret_cast<const float4*>(pa + vi * 8 + 4));
                v[i][0] += a.x; v[i][1] += a.y; v[i][2] += a.z; v[i][3] += a.w;
                v[i][4] += b.x; v[i][5] += b.y; v[i][6] += b.z; v[i][7] += b.w;
            }
        }
#pragma unroll
        for (int i = 0; i < MAX_VEC; ++i)
#pragma unroll
            for (int u = 0; u < 8; ++u) sum += v[i][u];
    } else
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = tr + i * TPR;
        if (active && vi < nvec) {
            if constexpr (IN_BF16) {
                const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + row * ld_x + vi * 8);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 f = __bfloat1622float2(h[u]);
                    v[i][2 * u] = f.x; v[i][2 * u + 1] = f.y;
                }
            } else {
                const float* px = reinterpret_cast<const float*>(x) + row * ld_x + vi * 8;
                const float4 a = *reinterpret_cast<const float4*>(px);
                const float4 b = *reinterpret_cast<const float4*>(px + 4);
                v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
                v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
            }
            if (pre_add != nullptr) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(pre_add + vi * 8));
                const float4 b = __ldg(reinterpret_cast<const float4*>(pre_add + vi * 8 + 4));
                v[i][0] += a.x; v[i][1] += a.y; v[i][2] += a.z; v[i][3] += a.w;
                v[i][4] += b.x; v[i][5] += b.y; v[i][6] += b.z; v[i][7] += b.w;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) sum += v[i][u];
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[i][u] = 0.f;
        }
    }
    auto row_reduce = [&](float val, int which) -> float {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if constexpr (TPR > 32) {
            if ((threadIdx.x & 31) == 0) red[which][threadIdx.x >> 5] = val;
            __syncthreads();
            val = 0.f;
#pragma unroll
            for (int w = 0; w < TPR / 32; ++w) val += red[which][w];
        }
        return val;
    };
    const float mean = row_reduce(sum, 0) / static_cast<float>(n);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = tr + i * TPR;
        if (EXACT || (active && vi < nvec)) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { const float d = v[i][u] - mean; sq += d * d; }
        }
    }
    const float var = row_reduce(sq, 1) / static_cast<float>(n);
    const float rstd = 1.0f / sqrtf(var + eps);
    if (!active) return;

    long long orow = row;
    if (grp_rows > 0) {
        const int g = row / grp_rows;
        orow = static_cast<long long>(g) * grp_stride + grp_off + (row - g * grp_rows);
    }
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = tr + i * TPR;
        if (EXACT || vi < nvec) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float y[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) y[u] = (v[i][u] - mean) * rstd * g[u] + bb[u];
            if constexpr (OUT_F32) {
                float* o = reinterpret_cast<float*>(out) + orow * ld_out + vi * 8;
                *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
            } else {
                uint4 q;
                q.x = pack_bf16(y[0], y[1]); q.y = pack_bf16(y[2], y[3]);
                q.z = pack_bf16(y[4], y[5]); q.w = pack_bf16(y[6], y[7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * ld_out + vi * 8) = q;
            }
        }
    }
}

// xb = bf16(x) and per-row (sum, sumsq) of xb: one warp per row, 16-byte accesses.
__global__ void __launch_bounds__(256)
rowstats_cast_kernel(const float* __restrict__ x, long long ld_x, __nv_bfloat16* __restrict__ xb, long long ld_xb,
                     float2* __restrict__ stats, int rows, int n) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* px = x + row * ld_x;
    __nv_bfloat16* po = xb + row * ld_xb;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane * 8; i < n; i += 256) {
        const float4 a = *reinterpret_cast<const float4*>(px + i);
        const float4 b = *reinterpret_cast<const float4*>(px + i + 4);
        uint4 q;
        q.x = pack_bf16(a.x, a.y); q.y = pack_bf16(a.z, a.w); q.z = pack_bf16(b.x, b.y); q.w = pack_bf16(b.z, b.w);
        *reinterpret_cast<uint4*>(po + i) = q;
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 r = __bfloat1622float2(h[u]);
            s1 += r.x + r.y;
            s2 = fmaf(r.x, r.x, fmaf(r.y, r.y, s2));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) stats[row] = make_float2(s1, s2);
}

// x[b, t, :] = in[b, t, :] + pos[t + 2, :]   (Decoder.forward_embedding with token_embedding given)
__global__ void __launch_bounds__(256)
add_positions_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int dim, const float* __restrict__ pos) {
    const int t = blockIdx.x % T;
    const float4* a = reinterpret_cast<const float4*>(in + static_cast<long long>(blockIdx.x) * dim);
    const float4* p = reinterpret_cast<const float4*>(pos + static_cast<long long>(t + 2) * dim);
    float4* o = reinterpret_cast<float4*>(out + static_cast<long long>(blockIdx.x) * dim);
    for (int i = threadIdx.x; i < dim / 4; i += blockDim.x) {
        const float4 x = a[i], c = __ldg(p + i);
        o[i] = make_float4(x.x + c.x, x.y + c.y, x.z + c.z, x.w + c.w);
    }
}

// ----------------------------------------------------------------------------- embedding + splice + positions
struct SpliceRows { int count; int start[KX_MAX_IMAGES]; };   // first spliced row of every image, ascending
__global__ void __launch_bounds__(256)
embed_splice_pos_kernel(const long long* __restrict__ tokens, int t_text, const float* __restrict__ embed, int vocab,
                        const float* __restrict__ pos, int dim, const SpliceRows img, int n_img, int alias, float* __restrict__ x0,
                        int* __restrict__ err_flag) {
    const int T = t_text + n_img * img.count;
    const int b = blockIdx.x / T;
    const int t = blockIdx.x - b * T;
    int ti = t;                                               // text index = row - 64 * (images that start before it)
    for (int i = 0; i < img.count; ++i) {
        if (t >= img.start[i]) {
            if (t < img.start[i] + n_img) return;             // image row: written by the image_proj GEMM epilogue
            ti -= n_img;
        }
    }
    long long tok = tokens[static_cast<long long>(b) * t_text + ti];
    if (tok < 0 || tok >= vocab) {
        if (err_flag != nullptr && threadIdx.x == 0) atomicExch(err_flag, 1);
        tok = 0;
    }
    const float4* e = reinterpret_cast<const float4*>(embed + tok * dim);
    float4* o = reinterpret_cast<float4*>(x0 + static_cast<long long>(blockIdx.x) * dim);
    if (pos == nullptr) {                                   // gather only (forward_embedding(...)[1])
        for (int i = threadIdx.x; i < dim / 4; i += blockDim.x) o[i] = __ldg(e + i);
        return;
    }
    const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<long long>(t + 2) * dim);
    if (alias) {
        // torchscale's `x = embed = scale * tok; x += positions` (in place): the embeddings taken at model.py:238 already
        // hold pos[ti + 2]; the second forward_embedding adds pos[t + 2].  Same association as the reference: (e + p1) + p2.
        const float4* p1 = reinterpret_cast<const float4*>(pos + static_cast<long long>(ti + 2) * dim);
        for (int i = threadIdx.x; i < dim / 4; i += blockDim.x) {
            const float4 a = __ldg(e + i), c1 = __ldg(p1 + i), c = __ldg(pp + i);
            o[i] = make_float4((a.x + c1.x) + c.x, (a.y + c1.y) + c.y, (a.z + c1.z) + c.z, (a.w + c1.w) + c.w);
        }
        return;
    }
    for (int i = threadIdx.x; i < dim / 4; i += blockDim.x) {
        const float4 a = __ldg(e + i), c = __ldg(pp + i);
        o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
}

// ----------------------------------------------------------------------------- CLIP pixel preprocessing on device
// CLIPImageProcessor's rescale + normalise for images that already have the model's size (SURVEY.md §8(f)4;
// reference call site kosmosx/model.py:81-97; HF 4.35 image_transforms.py rescale(): uint8 * (1/255) in float64,
// rounded to float32; normalize(): (x - mean) / std in float32).  Every step is written with explicit
// round-to-nearest intrinsics so that no FMA contraction changes the reference's roundings.
struct ClipNorm {
    float mean[3];
    float std[3];
};

// kernel parameters live in the constant bank: select instead of indexing dynamically (no local-memory copy)
__device__ __forceinline__ float pick3(const float (&a)[3], int c) { return c == 0 ? a[0] : (c == 1 ? a[1] : a[2]); }

__device__ __forceinline__ float clip_pixel(unsigned char u, float mean, float std) {
    const float r = __double2float_rn(__dmul_rn(static_cast<double>(u), 1.0 / 255.0));
    return __fdiv_rn(__fsub_rn(r, mean), std);
}

// 3 x 256 table of clip_pixel for one CTA: every later conversion is a shared-memory lookup of the same values
// (the float64 multiply and the IEEE divide would otherwise bound these kernels, not HBM)
__device__ __forceinline__ void build_clip_lut(float* lut, const ClipNorm& nm) {
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8;
        lut[i] = clip_pixel(static_cast<unsigned char>(i & 255), pick3(nm.mean, c), pick3(nm.std, c));
    }
    __syncthreads();
}

__device__ __forceinline__ long long u8_index(int channels_last, long long b, int c, int y, int x, int image) {
    return channels_last ? ((b * image + y) * image + x) * 3 + c : ((b * 3 + c) * image + y) * image + x;
}

// pixel_values (N,3,H,W) fp32 from uint8 (N,3,H,W) or (N,H,W,3): one thread per 4 consecutive output pixels
// (16-byte stores; the uint8 loads of a warp cover 128 consecutive bytes in the planar layout).
__global__ void __launch_bounds__(256)
clip_normalize_u8_kernel(const unsigned char* __restrict__ pixels, int channels_last, long long n_quads, int image, ClipNorm nm,
                         float* __restrict__ out) {
    __shared__ float lut[768];
    build_clip_lut(lut, nm);
    const int plane4 = image * image / 4;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_quads; i += gridDim.x * 256ll) {
        const long long bc = i / plane4;
        const int r = static_cast<int>(i - bc * plane4) * 4;
        const long long b = bc / 3;
        const int c = static_cast<int>(bc - b * 3);
        const int y = r / image, x = r - y * image;            // image % 4 == 0: the 4 pixels share a row
        const float* t = lut + (c << 8);
        float4 v;
        if (channels_last) {
            const unsigned char* p = pixels + u8_index(1, b, c, y, x, image);
            v = make_float4(t[p[0]], t[p[3]], t[p[6]], t[p[9]]);
        } else {
            const uchar4 q = *reinterpret_cast<const uchar4*>(pixels + u8_index(0, b, c, y, x, image));
            v = make_float4(t[q.x], t[q.y], t[q.z], t[q.w]);
        }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// ----------------------------------------------------------------------------- CLIP patch im2col + CLS rows
// Patch pack for the conv-as-GEMM patch embedding ([HF]:202-218), one CTA per strip of `patch` pixel rows of one image
// (image/patch patches): the strip is read once with coalesced 16-byte (fp32) / 4-byte (uint8) loads into shared memory
// as fp32 [c][dy][x] (uint8 through a per-CTA 3 x 256 table of the normalised values); then every patch row (k = c*p*p + dy*p + dx, zero padded to k_pad) is written as 16-byte chunks of
// 8 bf16, consecutive threads on consecutive chunks.  The k -> strip offset map is built once per CTA (no per-element
// divisions).  MODE 0: fp32 pixel_values (N,3,H,W).  MODE 1 / 2: raw uint8 pixels, planar / channels-last, with the
// rescale + normalise above fused in - the fp32 pixel_values tensor never exists in HBM.
// Blocks past the strips write the CLS rows x[b, 0, :] = class_embedding + pos[0].
template <int MODE>
__global__ void __launch_bounds__(256)
im2col_strip_kernel(const void* __restrict__ pixels_, ClipNorm nm, int batch, int media, int image, int patch,
                    __nv_bfloat16* __restrict__ patches, int k_pad, const float* __restrict__ cls,
                    const float* __restrict__ pos, float* __restrict__ x, int dim) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = image / patch;
    const int n_strips = batch * g;
    if (static_cast<int>(blockIdx.x) >= n_strips) {
        const int b = blockIdx.x - n_strips;
        float* o = x + static_cast<long long>(b) * (g * g + 1) * dim;
        for (int i = threadIdx.x; i < dim; i += blockDim.x) o[i] = cls[i] + pos[i];
        return;
    }
    const int strip_elems = 3 * patch * image;                // fp32 strip [3][patch][image]
    float* strip = reinterpret_cast<float*>(smem_raw);
    int* kmap = reinterpret_cast<int*>(strip + strip_elems);  // k -> offset of (c, dy, dx) inside the strip, -1 = padding
    float* lut = reinterpret_cast<float*>(kmap + k_pad);       // MODE 1 / 2 only
    if constexpr (MODE != 0) build_clip_lut(lut, nm);
    const int slot = blockIdx.x / g;                          // output image slot, media-major: slot = i * (batch/media) + seq
    const int py = blockIdx.x - slot * g;
    const int seqs = batch / media;
    const long long b = (slot % seqs) * media + slot / seqs;  // source image: pixels are (seq, media, ...)
    const int pp = patch * patch;
    for (int k = threadIdx.x; k < k_pad; k += blockDim.x) {
        int off = -1;
        if (k < 3 * pp) {
            const int c = k / pp, r = k - c * pp;
            const int dy = r / patch, dx = r - dy * patch;
            off = (c * patch + dy) * image + dx;
        }
        kmap[k] = off;
    }
    const int row4 = image >> 2;                              // image % 4 == 0 (checked on the host)
    if constexpr (MODE == 0) {
        const float* px = reinterpret_cast<const float*>(pixels_);
        for (int i = threadIdx.x; i < 3 * patch * row4; i += blockDim.x) {
            const int r = i / row4, q = i - r * row4;         // r = c * patch + dy
            const int c = r / patch, dy = r - c * patch;
            const float4 v = __ldg(reinterpret_cast<const float4*>(px + ((b * 3 + c) * image + py * patch + dy) * image) + q);
            reinterpret_cast<float4*>(strip)[i] = v;
        }
    } else if constexpr (MODE == 1) {
        const unsigned char* px = reinterpret_cast<const unsigned char*>(pixels_);
        for (int i = threadIdx.x; i < 3 * patch * row4; i += blockDim.x) {
            const int r = i / row4, q = i - r * row4;
            const int c = r / patch, dy = r - c * patch;
            const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(px + ((b * 3 + c) * image + py * patch + dy) * image) + q);
            const float* t = lut + (c << 8);
            reinterpret_cast<float4*>(strip)[i] = make_float4(t[u.x], t[u.y], t[u.z], t[u.w]);
        }
    } else {
        // channels-last: the strip is one contiguous run of patch * image * 3 bytes (a multiple of 4)
        const unsigned char* px = reinterpret_cast<const unsigned char*>(pixels_) + (b * image + py * patch) * image * 3;
        const int row_bytes = image * 3;
        for (int w = threadIdx.x; w < (patch * row_bytes) >> 2; w += blockDim.x) {
            const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(px) + w);
            const unsigned char bytes[4] = {u.x, u.y, u.z, u.w};
            int j = w << 2;
            int dy = j / row_bytes;
            int rem = j - dy * row_bytes;
            int xx = rem / 3, c = rem - xx * 3;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                strip[(c * patch + dy) * image + xx] = lut[(c << 8) + bytes[e]];
                if (++c == 3) { c = 0; if (++xx == image) { xx = 0; ++dy; } }
            }
        }
    }
    __syncthreads();
    const int chunks = k_pad >> 3;
    __nv_bfloat16* out = patches + (static_cast<long long>(slot) * g * g + static_cast<long long>(py) * g) * k_pad;
    for (int i = threadIdx.x; i < g * chunks; i += blockDim.x) {
        const int pxi = i / chunks, j = i - pxi * chunks;
        const float* src = strip + pxi * patch;
        const int* km = kmap + (j << 3);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int off = km[e];
            v[e] = off >= 0 ? src[off] : 0.f;
        }
        uint4 o;
        o.x = pack_bf16(v[0], v[1]);
        o.y = pack_bf16(v[2], v[3]);
        o.z = pack_bf16(v[4], v[5]);
        o.w = pack_bf16(v[6], v[7]);
        reinterpret_cast<uint4*>(out + static_cast<long long>(pxi) * k_pad)[j] = o;
    }
}

// ----------------------------------------------------------------------------- xPos tables
__global__ void xpos_tables_kernel(const float* __restrict__ scale, const float* __restrict__ inv_freq, int T,
                                   int min_pos, float scale_base, float* q_cos, float* q_sin, float* k_cos,
                                   float* k_sin) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * 32) return;
    const int t = idx >> 5, j = idx & 31;
    // the reference computes the exponent, the power and the angle in fp32; keep
    // those roundings, evaluate pow / sin / cos in double so the table is the correctly
    // rounded value of the reference's own fp32 arguments.
    const float expo = static_cast<float>(t + min_pos) / scale_base;
    const float S = static_cast<float>(pow(static_cast<double>(scale[j]), static_cast<double>(expo)));
    const float theta = static_cast<float>(t) * inv_freq[j];
    const float sn = static_cast<float>(sin(static_cast<double>(theta)));
    const float cs = static_cast<float>(cos(static_cast<double>(theta)));
    const float Sinv = 1.0f / S;
    q_cos[idx] = cs * S;
    q_sin[idx] = sn * S;
    k_cos[idx] = cs * Sinv;
    k_sin[idx] = sn * Sinv;
}

// ----------------------------------------------------------------------------- staging helpers
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long nv = n >> 3;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
        uint4 q;
        q.x = pack_bf16(a.x, a.y); q.y = pack_bf16(a.z, a.w);
        q.z = pack_bf16(b.x, b.y); q.w = pack_bf16(b.z, b.w);
        reinterpret_cast<uint4*>(dst)[i] = q;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const long long i = (nv << 3) + threadIdx.x;
        dst[i] = __float2bfloat16_rn(src[i]);
    }
}

__global__ void __launch_bounds__(256)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
    const long long nv = n >> 3;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float4 a, b;
        a.x = __uint_as_float(w[0] << 16); a.y = __uint_as_float(w[0] & 0xffff0000u);
        a.z = __uint_as_float(w[1] << 16); a.w = __uint_as_float(w[1] & 0xffff0000u);
        b.x = __uint_as_float(w[2] << 16); b.y = __uint_as_float(w[2] & 0xffff0000u);
        b.z = __uint_as_float(w[3] << 16); b.w = __uint_as_float(w[3] & 0xffff0000u);
        reinterpret_cast<float4*>(dst)[2 * i] = a;
        reinterpret_cast<float4*>(dst)[2 * i + 1] = b;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const long long i = (nv << 3) + threadIdx.x;
        dst[i] = __bfloat162float(src[i]);
    }
}

__global__ void __launch_bounds__(256)
broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long row_elems, int copies) {
    const long long total = row_elems * copies;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride)
        dst[i] = __ldg(src + i % row_elems);
}

}  // namespace kx

using namespace kx;

extern "C" int kx_layernorm_fwd(const void* x, int x_is_bf16, long long ld_x, const float* pre_add, int pre_add_group,
                                int pre_add_rows, const float* gamma, const float* beta, float eps, void* out,
                                int out_is_f32, long long ld_out, int rows, int n, int grp_rows, int grp_stride,
                                int grp_off, cudaStream_t stream) {
    if (!x || !gamma || !beta || !out) { set_error("kx_layernorm_fwd: null pointer"); return KX_ERR_ARG; }
    if (rows <= 0 || n <= 0 || (n % 8) || n > 32768) { set_error("kx_layernorm_fwd: n=%d must be a multiple of 8 and <= 32768", n); return KX_ERR_ARG; }
    const int in_align = x_is_bf16 ? 8 : 4;
    if (pre_add && pre_add_group > 0 && pre_add_rows <= 0) { set_error("kx_layernorm_fwd: pre_add_rows must be > 0"); return KX_ERR_ARG; }
    if ((ld_x % in_align) || (ld_out % (out_is_f32 ? 4 : 8)) || ((uintptr_t)x & 15) || ((uintptr_t)out & 15) || ((uintptr_t)gamma & 15) ||
        ((uintptr_t)beta & 15) || (pre_add && ((uintptr_t)pre_add & 15))) {
        set_error("kx_layernorm_fwd: pointers and row pitches must be 16-byte aligned");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
#define KX_LN3(BF, OF, TPR, MV, EX)                                                                                    \
    layernorm_kernel<BF, OF, TPR, MV, EX><<<(rows + (256 / TPR) - 1) / (256 / TPR), 256, 0, stream>>>(                  \
        x, ld_x, pre_add, pre_add_group, pre_add_rows, gamma, beta, eps, out, ld_out, rows, n, grp_rows, grp_stride, grp_off)
#define KX_LN2(BF, OF, TPR, MV)                                                                                        \
    do { if (n == (TPR) * (MV) * 8) KX_LN3(BF, OF, TPR, MV, true); else KX_LN3(BF, OF, TPR, MV, false); } while (0)
#define KX_LN(TPR, MV)                                                                                                 \
    do {                                                                                                               \
        if (x_is_bf16) { if (out_is_f32) KX_LN2(true, true, TPR, MV); else KX_LN2(true, false, TPR, MV); }              \
        else { if (out_is_f32) KX_LN2(false, true, TPR, MV); else KX_LN2(false, false, TPR, MV); }                      \
    } while (0)
    if (n <= 1024) KX_LN(32, 4);
    else if (n <= 2048) KX_LN(32, 8);
    else if (n <= 8192) KX_LN(256, 4);
    else KX_LN(256, 16);
#undef KX_LN
#undef KX_LN2
#undef KX_LN3
    return check_launch("kx_layernorm_fwd");
}

extern "C" int kx_rowstats_cast(const float* x, long long ld_x, void* xb, long long ld_xb, float* stats, int rows, int n,
                                cudaStream_t stream) {
    if (!x || !xb || !stats) { set_error("kx_rowstats_cast: null pointer"); return KX_ERR_ARG; }
    if (rows <= 0 || n <= 0 || (n % 8) || (ld_x % 4) || (ld_xb % 8) || ((uintptr_t)x & 15) || ((uintptr_t)xb & 15) ||
        ((uintptr_t)stats & 7)) {
        set_error("kx_rowstats_cast: n must be a multiple of 8 and rows 16-byte aligned");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    rowstats_cast_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, ld_x, reinterpret_cast<__nv_bfloat16*>(xb), ld_xb,
                                                            reinterpret_cast<float2*>(stats), rows, n);
    return check_launch("kx_rowstats_cast");
}

extern "C" int kx_embed_splice_pos(const long long* tokens, int batch, int t_text, const float* embed_table, int vocab,
                                   const float* pos_table, int pos_rows, int dim, const int* host_img_rows, int img_count,
                                   int n_img, int alias_positions, float* x0, int* err_flag, cudaStream_t stream) {
    if (!tokens || !embed_table || !x0) { set_error("kx_embed_splice_pos: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || t_text <= 0 || n_img < 0 || (dim % 4) || img_count < 0 || img_count > KX_MAX_IMAGES ||
        (img_count > 0 && !host_img_rows)) {
        set_error("kx_embed_splice_pos: bad shape (batch=%d t_text=%d n_img=%d dim=%d img_count=%d)", batch, t_text, n_img, dim, img_count);
        return KX_ERR_ARG;
    }
    const int T = t_text + n_img * img_count;
    SpliceRows img = {};
    img.count = img_count;
    for (int i = 0; i < img_count; ++i) {
        const int s = host_img_rows[i];
        // image i sits in front of text token s - i*n_img: blocks ascend, do not overlap, stay inside the sequence
        if (s < 0 || s + n_img > T || (i > 0 && s < host_img_rows[i - 1] + n_img)) {
            set_error("kx_embed_splice_pos: image %d cannot start at spliced row %d (T=%d, %d rows per image)", i, s, T, n_img);
            return KX_ERR_ARG;
        }
        img.start[i] = s;
    }
    if (pos_table && T + 2 > pos_rows) {
        set_error("kx_embed_splice_pos: sequence length %d needs %d position rows, table has %d", T, T + 2, pos_rows);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    embed_splice_pos_kernel<<<batch * T, 256, 0, stream>>>(tokens, t_text, embed_table, vocab, pos_table, dim, img, n_img,
                                                           (alias_positions && pos_table) ? 1 : 0, x0, err_flag);
    return check_launch("kx_embed_splice_pos");
}

extern "C" int kx_add_positions(const float* in, float* out, int batch, int T, int dim, const float* pos_table,
                                int pos_rows, cudaStream_t stream) {
    if (!in || !out || !pos_table || batch <= 0 || T <= 0 || (dim % 4)) { set_error("kx_add_positions: bad argument"); return KX_ERR_ARG; }
    if (T + 2 > pos_rows) {
        set_error("kx_add_positions: sequence length %d needs %d position rows, table has %d", T, T + 2, pos_rows);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    add_positions_kernel<<<batch * T, 256, 0, stream>>>(in, out, T, dim, pos_table);
    return check_launch("kx_add_positions");
}

static bool clip_norm_from_host(const float* mean3, const float* std3, ClipNorm* nm, const char* what) {
    if (!mean3 || !std3) { set_error("%s: null mean / std", what); return false; }
    for (int c = 0; c < 3; ++c) {
        if (!(std3[c] > 0.f)) { set_error("%s: std[%d] must be positive", what, c); return false; }
        nm->mean[c] = mean3[c];
        nm->std[c] = std3[c];
    }
    return true;
}

extern "C" int kx_clip_normalize_u8(const unsigned char* pixels, int channels_last, int batch, int image, const float* mean3,
                                    const float* std3, float* pixel_values, cudaStream_t stream) {
    if (!pixels || !pixel_values) { set_error("kx_clip_normalize_u8: null pointer"); return KX_ERR_ARG; }
    if (batch <= 0 || image <= 0 || (image % 4) || (reinterpret_cast<uintptr_t>(pixels) & 3) ||
        (reinterpret_cast<uintptr_t>(pixel_values) & 15)) {
        set_error("kx_clip_normalize_u8: bad shape or alignment (batch=%d image=%d; image %% 4 == 0, 4-byte aligned input, "
                  "16-byte aligned output)", batch, image);
        return KX_ERR_ARG;
    }
    ClipNorm nm;
    if (!clip_norm_from_host(mean3, std3, &nm, "kx_clip_normalize_u8")) return KX_ERR_ARG;
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const long long n_quads = static_cast<long long>(batch) * 3 * image * image / 4;
    const long long blocks = std::min<long long>((n_quads + 255) / 256, static_cast<long long>(sms) * 8);
    clip_normalize_u8_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pixels, channels_last != 0, n_quads, image, nm,
                                                                               pixel_values);
    return check_launch("kx_clip_normalize_u8");
}

// shared launch of the strip kernel: mode 0 fp32 planar, 1 uint8 planar, 2 uint8 channels-last
static int launch_im2col(const char* what, int mode, const void* pixels, const ClipNorm& nm, int batch, int media, int image,
                         int patch, void* patches_bf16, int k_pad, const float* class_embedding, const float* pos_table,
                         float* x, int dim, cudaStream_t stream) {
    if (!pixels || !patches_bf16 || !class_embedding || !pos_table || !x) { set_error("%s: null pointer", what); return KX_ERR_ARG; }
    if (batch <= 0 || media <= 0 || batch % media || patch <= 0 || image % patch || (image % 4) || k_pad < 3 * patch * patch ||
        (k_pad % 8)) {
        set_error("%s: bad shape (image=%d patch=%d k_pad=%d; image %% patch == 0, image %% 4 == 0, k_pad %% 8 == 0)", what, image,
                  patch, k_pad);
        return KX_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(pixels) & (mode == 0 ? 15 : 3)) || (reinterpret_cast<uintptr_t>(patches_bf16) & 15)) {
        set_error("%s: pixels must be %d-byte aligned and the patch rows 16-byte aligned", what, mode == 0 ? 16 : 4);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    const size_t smem = static_cast<size_t>(3) * patch * image * sizeof(float) + static_cast<size_t>(k_pad) * sizeof(int) +
                        (mode != 0 ? 768 * sizeof(float) : 0);
    if (smem > 200 * 1024) { set_error("%s: a %d-row strip of a %d-wide image does not fit in shared memory", what, patch, image); return KX_ERR_ARG; }
    const int g = image / patch;
    const unsigned blocks = static_cast<unsigned>(batch * g + batch);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(patches_bf16);
    cudaError_t e = cudaSuccess;
#define KX_IM2COL(MODE_)                                                                                                     \
    {                                                                                                                        \
        if (smem > 48 * 1024)                                                                                                \
            e = cudaFuncSetAttribute(im2col_strip_kernel<MODE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        if (e == cudaSuccess)                                                                                                \
            im2col_strip_kernel<MODE_><<<blocks, 256, smem, stream>>>(pixels, nm, batch, media, image, patch, out, k_pad,    \
                                                                      class_embedding, pos_table, x, dim);                   \
    }
    if (mode == 0) KX_IM2COL(0) else if (mode == 1) KX_IM2COL(1) else KX_IM2COL(2)
#undef KX_IM2COL
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e)); return KX_ERR_LAUNCH; }
    return check_launch(what);
}

extern "C" int kx_im2col_patches(const float* pixels, int batch, int media, int image, int patch, void* patches_bf16, int k_pad,
                                 const float* class_embedding, const float* pos_table, float* x, int dim,
                                 cudaStream_t stream) {
    ClipNorm nm = {{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
    return launch_im2col("kx_im2col_patches", 0, pixels, nm, batch, media, image, patch, patches_bf16, k_pad, class_embedding,
                         pos_table, x, dim, stream);
}

extern "C" int kx_im2col_patches_u8(const unsigned char* pixels, int channels_last, const float* mean3, const float* std3,
                                    int batch, int media, int image, int patch, void* patches_bf16, int k_pad,
                                    const float* class_embedding, const float* pos_table, float* x, int dim,
                                    cudaStream_t stream) {
    ClipNorm nm;
    if (!clip_norm_from_host(mean3, std3, &nm, "kx_im2col_patches_u8")) return KX_ERR_ARG;
    return launch_im2col("kx_im2col_patches_u8", channels_last ? 2 : 1, pixels, nm, batch, media, image, patch, patches_bf16, k_pad,
                         class_embedding, pos_table, x, dim, stream);
}

extern "C" int kx_xpos_tables(const float* scale, const float* inv_freq, int T, int min_pos, float scale_base,
                              float* q_cos, float* q_sin, float* k_cos, float* k_sin, cudaStream_t stream) {
    if (!scale || !inv_freq || !q_cos || !q_sin || !k_cos || !k_sin || T <= 0) { set_error("kx_xpos_tables: bad argument"); return KX_ERR_ARG; }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    xpos_tables_kernel<<<(T * 32 + 255) / 256, 256, 0, stream>>>(scale, inv_freq, T, min_pos, scale_base, q_cos, q_sin, k_cos, k_sin);
    return check_launch("kx_xpos_tables");
}

extern "C" int kx_cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
    if (!src || !dst || n <= 0 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) { set_error("kx_cast_f32_to_bf16: bad argument"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const long long nv = (n >> 3) + 1;
    const int blocks = static_cast<int>(std::min<long long>((nv + 255) / 256, static_cast<long long>(sms) * 8));
    cast_f32_bf16_kernel<<<blocks, 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    return check_launch("kx_cast_f32_to_bf16");
}

extern "C" int kx_cast_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t stream) {
    if (!src || !dst || n <= 0 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) {
        set_error("kx_cast_bf16_to_f32: null / unaligned pointer or n <= 0");
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    const long long blocks = std::min<long long>(((n >> 3) + 255) / 256 + 1, 148ll * 16);
    cast_bf16_f32_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
    return check_launch("kx_cast_bf16_to_f32");
}

extern "C" int kx_broadcast_rows(const float* src, float* dst, long long row_elems, int copies, cudaStream_t stream) {
    if (!src || !dst || row_elems <= 0 || copies <= 0) { set_error("kx_broadcast_rows: bad argument"); return KX_ERR_ARG; }
    const int sms = device_sm_count();
    if (sms <= 0) return KX_ERR_NO_DEVICE;
    const long long total = row_elems * copies;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sms) * 8));
    broadcast_rows_kernel<<<blocks, 256, 0, stream>>>(src, dst, row_elems, copies);
    return check_launch("kx_broadcast_rows");
}
