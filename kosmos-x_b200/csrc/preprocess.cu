// Host preprocessing moved onto the device: CLIPImageProcessor's resize (shortest edge -> 224, PIL
// bicubic) and centre crop, which KosmosTokenizer.tokenize_images applies on the host through the processor built at
// /root/reference/kosmosx/model.py:36-38 and called at model.py:81-97 (transformers 4.35 slow processor:
// image_transforms.resize -> PIL.Image.resize(..., BICUBIC), then center_crop).
//
// PIL resamples 8-bit images in fixed point (ImagingResample, 8bpc path): a horizontal pass then a vertical pass, each a
// 1-D convolution with per-output-pixel coefficient rows kk[xx][0..xmax) (doubles normalised to sum 1, scaled by 2^22 and
// rounded half away from zero) starting at input index xmin; the accumulator starts at 2^21, the result is
// clip8(acc >> 22), and the image between the passes is uint8.  The two kernels below are exactly that integer
// arithmetic, so the result is BIT-IDENTICAL to PIL's given the same coefficient tables (built on the host in float64
// by kosmosx/preprocess.py, a few KB per image size).  The centre crop is fused: the tables hold only the output columns
// / rows inside the crop window, and the horizontal pass only the input rows the vertical pass will read.
#include "kx_internal.h"

namespace kx {

constexpr int RS_BITS = 32 - 8 - 2;      // PIL's PRECISION_BITS

__device__ __forceinline__ unsigned char rs_clip8(int acc) {
    const int v = acc >> RS_BITS;
    return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// src: uint8, (n, in_h, in_w, 3) channels-last or (n, 3, in_h, in_w) planar.  tmp: (n, rows, out_w, 3) channels-last, rows =
// input rows [y0, y0 + rows).  One thread per (image, row, output column): the 3 channels share the coefficient row.
__global__ void __launch_bounds__(256)
resize_h_kernel(const unsigned char* __restrict__ src, int channels_last, int n, int in_h, int in_w, int y0, int rows,
                const int* __restrict__ kk, const int* __restrict__ bounds, int ksize, int out_w, unsigned char* __restrict__ tmp) {
    const long long total = static_cast<long long>(n) * rows * out_w;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int xx = static_cast<int>(i % out_w);
        const long long r = i / out_w;
        const int y = static_cast<int>(r % rows) + y0;
        const long long b = r / rows;
        const int xmin = __ldg(bounds + 2 * xx), xmax = __ldg(bounds + 2 * xx + 1);
        const int* k = kk + static_cast<long long>(xx) * ksize;
        int a0 = 1 << (RS_BITS - 1), a1 = a0, a2 = a0;
        if (channels_last) {
            const unsigned char* p = src + ((b * in_h + y) * in_w + xmin) * 3;
            for (int x = 0; x < xmax; ++x) {
                const int c = __ldg(k + x);
                a0 += p[3 * x] * c; a1 += p[3 * x + 1] * c; a2 += p[3 * x + 2] * c;
            }
        } else {
            const long long plane = static_cast<long long>(in_h) * in_w;
            const unsigned char* p = src + (b * 3) * plane + static_cast<long long>(y) * in_w + xmin;
            for (int x = 0; x < xmax; ++x) {
                const int c = __ldg(k + x);
                a0 += p[x] * c; a1 += p[plane + x] * c; a2 += p[2 * plane + x] * c;
            }
        }
        unsigned char* o = tmp + i * 3;
        o[0] = rs_clip8(a0); o[1] = rs_clip8(a1); o[2] = rs_clip8(a2);
    }
}

// tmp: (n, rows, out_w, 3) holding input rows [y0, ...); dst: (n, out_h, out_w, 3) channels-last.
__global__ void __launch_bounds__(256)
resize_v_kernel(const unsigned char* __restrict__ tmp, int n, int rows, int y0, const int* __restrict__ kk,
                const int* __restrict__ bounds, int ksize, int out_h, int out_w, unsigned char* __restrict__ dst) {
    const int row_bytes = out_w * 3;
    const long long total = static_cast<long long>(n) * out_h * row_bytes;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int xc = static_cast<int>(i % row_bytes);
        const long long r = i / row_bytes;
        const int yy = static_cast<int>(r % out_h);
        const long long b = r / out_h;
        const int ymin = __ldg(bounds + 2 * yy), ymax = __ldg(bounds + 2 * yy + 1);
        const int* k = kk + static_cast<long long>(yy) * ksize;
        const unsigned char* p = tmp + (b * rows + (ymin - y0)) * row_bytes + xc;
        int acc = 1 << (RS_BITS - 1);
        for (int y = 0; y < ymax; ++y) acc += p[static_cast<long long>(y) * row_bytes] * __ldg(k + y);
        dst[i] = rs_clip8(acc);
    }
}

}  // namespace kx

using namespace kx;

extern "C" int kx_resize_crop_u8(const unsigned char* pixels, int channels_last, int n, int in_h, int in_w, const int* kx,
                                 const int* bx, int ksize_x, const int* ky, const int* by, int ksize_y, int y0, int rows,
                                 int out_h, int out_w, unsigned char* tmp, unsigned char* out, cudaStream_t stream) {
    if (!pixels || !kx || !bx || !ky || !by || !tmp || !out) { set_error("kx_resize_crop_u8: null pointer"); return KX_ERR_ARG; }
    if (n <= 0 || in_h <= 0 || in_w <= 0 || ksize_x <= 0 || ksize_y <= 0 || out_h <= 0 || out_w <= 0 || y0 < 0 || rows <= 0 ||
        y0 + rows > in_h) {
        set_error("kx_resize_crop_u8: bad shape (n=%d in=%dx%d out=%dx%d rows [%d, %d))", n, in_h, in_w, out_h, out_w, y0, y0 + rows);
        return KX_ERR_ARG;
    }
    if (device_sm_count() <= 0) return KX_ERR_NO_DEVICE;
    auto grid = [](long long total) { return static_cast<int>(std::min<long long>((total + 255) / 256, 148ll * 32)); };
    resize_h_kernel<<<grid(static_cast<long long>(n) * rows * out_w), 256, 0, stream>>>(pixels, channels_last ? 1 : 0, n, in_h, in_w, y0,
                                                                                         rows, kx, bx, ksize_x, out_w, tmp);
    int st = check_launch("kx_resize_crop_u8 (horizontal)");
    if (st != KX_OK) return st;
    resize_v_kernel<<<grid(static_cast<long long>(n) * out_h * out_w * 3), 256, 0, stream>>>(tmp, n, rows, y0, ky, by, ksize_y, out_h,
                                                                                              out_w, out);
    return check_launch("kx_resize_crop_u8 (vertical)");
}
