/* kosmosx_b200.h — C ABI of libkosmosx_sm100.so
 *
 * The drop-in boundary of the B200-native Kosmos-X forward path.  The reference
 * (/root/reference/kosmosx/model.py) is a Python nn.Module; it has no FFI of its own, so
 * the entry points below are the operators its forward dispatches to through PyTorch
 * (F.linear, softmax/bmm attention, layer_norm, embedding/cat splice ...), re-cut at the
 * fusion boundaries of the sm_100a kernels.  Each declaration cites the reference call
 * site (file:line under /root/reference, or [HF] = transformers/models/clip/modeling_clip.py,
 * or SURVEY.md Appendix A for the un-vendored torchscale / flamingo_pytorch semantics).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - all kernels are enqueued on `stream`; nothing synchronises, nothing allocates;
 *   - return value: KX_OK (0) or a negative KX_ERR_* code; kx_last_error() gives the text;
 *   - bf16 = __nv_bfloat16 storage (uint16_t), row-major, `ld*` = row pitch in ELEMENTS.
 */
#ifndef KOSMOSX_B200_H
#define KOSMOSX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KX_ABI_VERSION 18   /* returned by kx_abi_version(); bumped on any signature change */

typedef struct CUstream_st* kx_stream_t; /* == cudaStream_t */

enum {
    KX_OK = 0,
    KX_ERR_ARG = -1,       /* bad argument (shape, alignment, null pointer) */
    KX_ERR_NO_DEVICE = -2, /* no CUDA device / driver: there is NO CPU fallback */
    KX_ERR_TMAP = -3,      /* cuTensorMapEncodeTiled rejected a tensor */
    KX_ERR_LAUNCH = -4     /* kernel launch failed */
};

enum { KX_ACT_NONE = 0, KX_ACT_GELU = 1 /* erf */, KX_ACT_QUICK_GELU = 2 /* x*sigmoid(1.702x) */ };
enum { KX_EPI_GENERIC = 0, KX_EPI_QKV_XPOS = 1 };
#define KX_MAX_IMAGES 16   /* images spliced into one sequence (kx_embed_splice_pos) */

/* ---- library ---------------------------------------------------------------------- */
const char* kx_last_error(void);      /* thread-local text of the last failure */
int kx_abi_version(void);             /* bumped on any signature change */
int kx_device_check(void);            /* KX_OK iff device 0.. current is sm_100 (B200) */
unsigned long long kx_launch_count(void);   /* kernels launched by this library so far */

/* ---- Linear layers ----------------------------------------------------------------- *
 * out = epilogue(A[M,K] . W[N,K]^T): tcgen05/TMEM GEMM with fused epilogue.
 * Replaces: HF CLIP q/k/v/out_proj, fc1/fc2 ([HF]:295-298,344-351) and the conv patch
 * embedding ([HF]:148-154,209, as GEMM over im2col rows); flamingo to_q/to_kv/to_out/FF
 * (SURVEY A.2); image_proj (model.py:205-206,232); torchscale q/k/v/out_proj, fc1, fc2
 *  and output_projection (model.py:166).
 *   epilogue order: +bias -> [xPos rotation of q,k column pairs] -> activation
 *                   -> +add_tab[(m % grp_rows) + add_off] -> +res[out_row] -> store(out_row)
 *   out_row = grp_rows ? (m / grp_rows) * grp_stride + grp_off + m % grp_rows : m
 */
typedef struct kx_gemm_args {
    int M, N, K;
    const float* bias;            /* [N] fp32 or NULL */
    const float* res;             /* fp32 [rows, ld_res] residual (may alias out) or NULL */
    long long ld_res;
    void* out;                    /* bf16 or fp32 [rows, ld_out] */
    long long ld_out;
    int out_f32;                  /* 1: fp32 output, 0: bf16 output */
    int act;                      /* KX_ACT_* */
    int epi;                      /* KX_EPI_* */
    int grp_rows, grp_stride, grp_off;    /* output row scatter (0 = identity) */
    const float* add_tab;         /* fp32 [*, ld_add] periodic additive table or NULL */
    int add_off;
    long long ld_add;
    const float *xq_cos, *xq_sin, *xk_cos, *xk_sin;  /* KX_EPI_QKV_XPOS: [seq_len,32] fp32 (kx_xpos_tables) */
    int seq_len;                  /* position of row m is m % seq_len */
    int d_model;                  /* N == 3*d_model: q | k | v column blocks */
    int cta_group;                /* 0 = auto, 1 = single CTA tiles, 2 = cta_group::2 pairs */
    int block_n;                  /* 0 = auto, 128 or 256 */
    int max_ctas;                 /* 0 = all SMs */
    int epi_mode;                 /* 0 = auto (staged TMA-store epilogue when alignment allows), 1 = direct stores */
    /* LayerNorm folded into this GEMM (consumer side).  A holds the RAW (un-normalised) rows,
     * W must already be W*diag(gamma), bias must be W.beta + b, and
     *     out = rstd[m] * (A.W^T - mean[m] * ln_c[n]) + bias[n] ...
     * mean/rstd of row m are computed from ln_tiles partial (sum, sumsq) pairs over ln_cols columns. */
    const float* ln_part;         /* fp32 pairs [ln_tiles][M][2], or NULL = no fold */
    const float* ln_c;            /* fp32 [N]: sum_k W[n,k] of the (bf16) folded weight */
    int ln_tiles, ln_cols;
    float ln_eps;
    /* producer side: per-row partial (sum, sumsq) of the values this GEMM stores, one pair per half tile
     * ([ceil(N/128)][M][2] fp32 with block_n = 256, the default here; [ceil(N/64)][M][2] with block_n = 128),
     * and a bf16 copy of an fp32 output; the
     * statistics are those of the bf16 copy when one is written.  Both need the staged epilogue. */
    float* stats_out;
    void* out2;
    long long ld_out2;
    /* operands given transposed in memory (the backward GEMMs read activations / weights where they lie):
     *   a_trans: A points at a [K, M] row-major matrix (lda >= M);  b_trans: W points at a [K, N] row-major matrix.
     *   dgrad  dX[M,Kd] = dY[M,N] . W[N,Kd]      -> A = dY, W = the forward weight with b_trans = 1
     *   wgrad  dW[N,Kd] = dY[M,N]^T . X[M,Kd]    -> A = dY with a_trans = 1, W = X with b_trans = 1
     * (replaces autograd's mm_backward of every F.linear on the path).  Generic epilogue only. */
    int a_trans, b_trans;
    /* training only: dropout with probability drop_p on the Linear's output, after bias / activation and BEFORE the residual
     * add — torchscale's `x = dropout(self_attn(...))` / `x = dropout(fc2(...))` (reference trains with dropout = 0.1,
     * kosmosx/model.py:175).  The mask is a pure function of (drop_seed, drop_site, row, column): Philox4x32-7, one call
     * per 8 consecutive columns, keep iff the 16-bit lot < round((1 - p) * 65536), kept values scaled by 65536 / that.
     * kx_layernorm_bwd regenerates it for the gradient (nothing is stored).  0 = off.  Plain epilogue only. */
    float drop_p;
    unsigned int drop_site;
    unsigned long long drop_seed;
} kx_gemm_args;

int kx_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const kx_gemm_args* args,
                 kx_stream_t stream);

/* ---- attention -------------------------------------------------------------------- *
 * Flash attention forward, head_dim 64, tcgen05 (S = Q.K^T and P.V on tensor cores, online
 * softmax in fp32).  q/k/v are column blocks of token-major matrices: row (b*T + t),
 * column (head*64 + d) relative to the given base pointer.
 * Replaces: torchscale MultiheadAttention core — bmm, nan_to_num, +triu(-inf) mask,
 * softmax(fp32), bmm, head merge (decoder, causal=1) — and HF CLIPAttention's
 * eager/sdpa core ([HF]:318-331; ViT, causal=0).  `scale` multiplies q.k^T (64^-0.5).
 * stats_out (fp32 pairs [heads][batch*seq_len][2], or NULL): per-head partial (sum, sumsq) of each stored
 * output row, consumed by the out_proj GEMM that folds torchscale's inner_attn_ln (kx_gemm_args.ln_part).
 */
int kx_attn_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out,
                int batch, int heads, int seq_len, int causal, float scale, float* stats_out, kx_stream_t stream);

/* Profiling aid for kx_attn_fwd (causal): with a device buffer of 4*64*8 int64 installed, CTA (0,0) of every
 * launch records clock64 stamps [role: softmax A, softmax B, MMA thread A, MMA thread B][KV block][point]; NULL = off. */
int kx_attn_set_trace(long long* device_buffer);

/* Perceiver cross-attention core (flamingo_pytorch PerceiverAttention; reached from
 * model.py:231): per (batch, head) softmax(q.k^T * scale - rowmax) . v with n_q latent queries and
 * n_kv = media + latent keys, head_dim 64.  q: [batch*n_q, ld_q], kv: [batch*n_kv, ld_kv] with
 * k at column head*64 and v at column v_col_off + head*64.  out: bf16 [batch*n_q, ld_out].
 * Runs on the tensor-core flash kernel of kx_attn_fwd (its keys-per-sequence != queries-per-sequence form). */
int kx_perceiver_xattn_fwd(const void* q, long long ld_q, const void* kv, long long ld_kv, int v_col_off, void* out,
                           long long ld_out, int batch, int heads, int n_q, int n_kv, float scale,
                           kx_stream_t stream);

/* ---- LayerNorm -------------------------------------------------------------------- *
 * y = (x + pre_add) normalised over the last dim, * gamma + beta, written as bf16 (GEMM
 * operand) at out_row (same scatter rule as kx_gemm_args).  x is fp32 (residual stream)
 * or bf16.  eps 1e-5 everywhere in the reference.  Replaces nn.LayerNorm at: [HF]:359,361,
 * 677 (ViT), flamingo norm_media/norm_latents/FF norm/final norm (A.2), torchscale
 * self_attn_layer_norm, inner_attn_ln, final_layer_norm, ffn_layernorm, decoder.layer_norm
 * (A.4).  pre_add ([pre_add_rows, N] fp32 or NULL) fuses `x + media_pos_emb[:m]` of the perceiver
 * (A.2): row (r / pre_add_group) % pre_add_rows is added to input row r (pre_add_group 0 = row 0 always).
 * out is bf16 (GEMM operand) or fp32 (ViT pre_layrnorm, whose output is the residual stream).
 */
int kx_layernorm_fwd(const void* x, int x_is_bf16, long long ld_x, const float* pre_add, int pre_add_group,
                     int pre_add_rows, const float* gamma, const float* beta, float eps, void* out, int out_is_f32,
                     long long ld_out, int rows, int n, int grp_rows, int grp_stride, int grp_off, kx_stream_t stream);

/* ---- embedding / splice ----------------------------------------------------------- *
 * x0[b, t, :] for every NON-image row of the spliced sequence (image rows are written by the
 * image_proj GEMM epilogue): token embedding gather + learned position (index t + 2).
 * The spliced sequence has T = t_text + img_count * n_img rows; image i occupies rows
 * [host_img_rows[i], host_img_rows[i] + n_img) (HOST array of img_count <= KX_MAX_IMAGES ascending,
 * non-overlapping starts) and the text tokens fill the remaining rows in order.
 * Replaces Decoder.forward_embedding x2 + torch.cat of model.py:238-244: the reference
 * is img_count = 1, host_img_rows = {2}; BASELINE.json configs[4] uses four images per sequence.
 * For KosmosLanguage (model.py:310-320) pass img_count = 0.
 * Token ids outside [0, vocab) set *err_flag (device int, may be NULL) instead of faulting.
 * pos_table NULL = gather only.
 * alias_positions = 1 reproduces torchscale's in-place add: forward_embedding does `x = embed = embed_scale *
 * token_embedding; x += positions`, so the `[1]` result taken at model.py:238 ALREADY carries positions 2..t_text+1 of
 * the un-spliced text, and the second call (model.py:242-244) adds the spliced positions on top: a text row holding
 * text token i at spliced row t gets embed[token] + pos[i + 2] + pos[t + 2] (image rows only pos[t + 2], from the
 * image_proj epilogue).  alias_positions = 0: a single add of pos[t + 2] (KosmosLanguage, model.py:310-320, which
 * calls forward_embedding once and takes `[0]`).
 */
int kx_embed_splice_pos(const long long* tokens, int batch, int t_text, const float* embed_table, int vocab,
                        const float* pos_table, int pos_rows, int dim, const int* host_img_rows, int img_count,
                        int n_img, int alias_positions, float* x0, int* err_flag, kx_stream_t stream);

/* x[b,t,:] = in[b,t,:] + pos_table[t+2,:]: Decoder.forward_embedding(x, token_embedding=x)[0] of
 * model.py:242-244 as a stand-alone call (the fused path is kx_embed_splice_pos + the image_proj epilogue). */
int kx_add_positions(const float* in, float* out, int batch, int T, int dim, const float* pos_table, int pos_rows,
                     kx_stream_t stream);

/* CLIP patch embedding front end ([HF]:202-218): im2col of (B,3,H,W) pixels (fp32) into bf16
 * rows [B*gh*gw, k_pad] (k = c*p*p + dy*p + dx, zero padded to k_pad), and the CLS rows
 * x[b, 0, :] = class_embedding + pos[0] written into the fp32 token buffer x [B, 1+gh*gw, dim].
 * media > 1: pixels are (B/media, media, 3, H, W) and output slot i*(B/media) + s holds image i of
 * sequence s (media-major), so that each image index is one contiguous row block downstream.
 * image % 4 == 0; pixels and patches_bf16 16-byte aligned (one CTA stages a strip of `patch` pixel rows in
 * shared memory with 16-byte loads and writes the patch rows as 16-byte chunks).
 */
int kx_im2col_patches(const float* pixels, int batch, int media, int image, int patch, void* patches_bf16, int k_pad,
                      const float* class_embedding, const float* pos_table, float* x, int dim, kx_stream_t stream);

/* Host preprocessing moved onto the device: CLIPImageProcessor's rescale + normalise, which
 * KosmosTokenizer.tokenize_images applies on the host (kosmosx/model.py:81-97), for uint8 images that already have
 * the model's size.  pixels: (N,3,H,W) (channels_last = 0) or (N,H,W,3) (channels_last = 1) uint8;
 * pixel_values[n,c,y,x] = (float32(float64(u8) * (1/255)) - mean[c]) / std[c] with float32 subtract / divide,
 * the roundings of HF 4.35 image_transforms.rescale / normalize.  mean3 / std3 are HOST arrays of 3 floats
 * (passed to the kernel by value).  image % 4 == 0. */
int kx_clip_normalize_u8(const unsigned char* pixels, int channels_last, int batch, int image, const float* mean3,
                         const float* std3, float* pixel_values, kx_stream_t stream);

/* kx_clip_normalize_u8 fused into kx_im2col_patches: uint8 pixels -> normalised bf16 patch rows + CLS rows; the
 * result is bit-identical to kx_im2col_patches applied to kx_clip_normalize_u8's output.  media and shape rules as
 * there; pixels 4-byte aligned. */
int kx_im2col_patches_u8(const unsigned char* pixels, int channels_last, const float* mean3, const float* std3, int batch,
                         int media, int image, int patch, void* patches_bf16, int k_pad, const float* class_embedding,
                         const float* pos_table, float* x, int dim, kx_stream_t stream);

/* CLIPImageProcessor's resize (shortest edge -> model size, PIL bicubic) + centre crop on the device for uint8 images of
 * any size (kosmosx/model.py:36-38,81-97 -> transformers 4.35 image_transforms.resize = PIL.Image.resize(BICUBIC), then
 * center_crop).  PIL's 8-bit resampling is fixed-point: horizontal pass, uint8 intermediate, vertical pass, each a 1-D
 * convolution acc = 2^21 + sum_x px[min + x] * k[x], out = clip8(acc >> 22).  kx / ky: DEVICE int32 coefficient rows
 * [out_w][ksize_x] / [out_h][ksize_y] (doubles normalised to 1, * 2^22, rounded half away from zero — built on the host,
 * kosmosx/preprocess.py) and bx / by: DEVICE int32 (first input index, tap count) pairs; only the output columns / rows
 * INSIDE the crop window are passed, so the crop costs nothing.  The horizontal pass computes input rows [y0, y0+rows)
 * (what the vertical taps read) into tmp (n * rows * out_w * 3 bytes); out: uint8 (n, out_h, out_w, 3) channels-last,
 * the layout kx_im2col_patches_u8 / kx_clip_normalize_u8 take with channels_last = 1.  Bit-identical to PIL. */
int kx_resize_crop_u8(const unsigned char* pixels, int channels_last, int n, int in_h, int in_w, const int* kx, const int* bx,
                      int ksize_x, const int* ky, const int* by, int ksize_y, int y0, int rows, int out_h, int out_w,
                      unsigned char* tmp, unsigned char* out, kx_stream_t stream);

/* Row statistics + bf16 copy of an fp32 matrix: xb = bf16(x), stats[m] = (sum, sumsq) of xb's row m
 * ([1][rows][2] fp32).  Seeds the folded-LayerNorm chain (kx_gemm_args.ln_part) for the first decoder layer,
 * whose input comes from kx_embed_splice_pos / the image_proj epilogue rather than from a GEMM with stats_out. */
int kx_rowstats_cast(const float* x, long long ld_x, void* xb_bf16, long long ld_xb, float* stats, int rows, int n,
                     kx_stream_t stream);

/* xPos tables (torchscale XPOS): for t in [0,T), j in [0,32):
 *   S = scale[j] ** ((t + min_pos) / scale_base), theta = t * inv_freq[j]
 *   q_cos = cos*S, q_sin = sin*S, k_cos = cos/S, k_sin = sin/S     (each [T,32] fp32)
 */
int kx_xpos_tables(const float* scale, const float* inv_freq, int T, int min_pos, float scale_base, float* q_cos,
                   float* q_sin, float* k_cos, float* k_sin, kx_stream_t stream);

/* fp32 -> bf16 conversion of a parameter tensor (weight staging), and broadcast of the
 * perceiver latents over the batch (A.2: latents.expand). */
int kx_cast_f32_to_bf16(const float* src, void* dst_bf16, long long n, kx_stream_t stream);
/* bf16 -> fp32 (16-byte aligned): the received sums of a bf16 gradient exchange go back into the flat fp32 gradient
 * buffer (2.76 GB of bf16 gradients per all-reduce; the reference's FSDP reduce_dtype is 16-bit,
 * train.py:156-162). */
int kx_cast_bf16_to_f32(const void* src_bf16, float* dst, long long n, kx_stream_t stream);
int kx_broadcast_rows(const float* src, float* dst, long long row_elems, int copies, kx_stream_t stream);

/* ---- verification precision ("bf16x3") ------------------------------------------------ *
 * BASELINE.json states the parity tolerance as logits max-abs-diff <= 1e-3 against the reference's PyTorch path
 * (kosmosx/model.py:208-253).  bf16 operands cannot reach it (one ulp at 1.0 is 7.8e-3), so the SAME tcgen05 GEMM is
 * also driven with split operands: x = hi + lo, hi = bf16(x), lo = bf16(x - hi), and A.W^T ~= Ah.Wh^T + Ah.Wl^T +
 * Al.Wh^T as ONE kx_gemm_bf16 launch over K' = 3*n_pad (three passes accumulating in one TMEM tile).
 * kx_split_bf16x3 writes that layout: dst[r, s*n_pad + k], s = 0,1,2 = (hi, hi, lo) for activations (weights = 0) or
 * (hi, lo, hi) for weights (weights = 1); columns k in [n, n_pad) are zero.  The remaining fp32 pieces of that mode:
 * kx_attn_f32 (softmax(q.k^T*scale [+causal]).v in fp32 FMAs, head_dim 64; q rows [batch*n_q, ld_q], k/v rows
 * [batch*n_kv, ld_kv], head h at columns h*64.. of the given pointers; serves the decoder (causal, n_q == n_kv), the ViT
 * and the perceiver cross-attention), kx_xpos_apply_f32 (the KX_EPI_QKV_XPOS rotation in place on the q|k blocks of an
 * fp32 [rows, ld] matrix) and kx_im2col_patches_f32 (kx_im2col_patches with fp32 patch rows). */
int kx_split_bf16x3(const float* src, long long ld_src, int rows, int n, int n_pad, void* dst_bf16, long long ld_dst,
                    int weights, kx_stream_t stream);
int kx_attn_f32(const float* q, long long ld_q, const float* k, const float* v, long long ld_kv, float* out,
                long long ld_out, int batch, int heads, int n_q, int n_kv, int causal, float scale, kx_stream_t stream);
int kx_xpos_apply_f32(float* qkv, long long ld, int rows, int d_model, int seq_len, const float* q_cos, const float* q_sin,
                      const float* k_cos, const float* k_sin, kx_stream_t stream);
int kx_im2col_patches_f32(const float* pixels, int batch, int media, int image, int patch, float* patches, int k_pad,
                          const float* class_embedding, const float* pos_table, float* x, int dim, kx_stream_t stream);

/* ==================================================================================== *
 * Training step (BASELINE.json configs[3]): forward -> cross-entropy over the text rows
 * -> backward -> gradient clipping -> AdamW / Lion.  The reference reaches these through autograd
 * (train.py:647-648 `loss = model(...)`, `accelerator.backward(loss)`), torch.nn.utils.clip_grad_norm_
 * (train.py:652-653) and torch.optim.AdamW / lion_pytorch.Lion (train.py:375-386).  The backward GEMMs are
 * kx_gemm_bf16 with a_trans / b_trans.
 * ==================================================================================== */

/* Attention backward (flash, head_dim 64, tcgen05): dq, dk, dv from d(out), q, k, v, out and the row log-sum-exp
 * written by kx_attn_fwd_lse (fp32 [heads][batch][ceil(seq_len/128)*128], log2 units).  Same layout rules as
 * kx_attn_fwd; dq/dk/dv are column blocks sharing ld_dqkv.  When the four xPos tables (kx_xpos_tables) are given,
 * dq and dk are returned as gradients of the UN-rotated projections (the transpose of the KX_EPI_QKV_XPOS rotation is
 * applied on the way out); NULL tables = plain attention.  Scratch: dq_accum fp32 [batch*seq_len, heads*64] (zeroed
 * by the call) and delta fp32 (TWICE the size of lse: (-lse, -rowsum(dO*O)) pairs).  Three kernels: delta = rowsum(dO*O) (which also
 * zeroes dq_accum), the main kernel (one CTA per (batch, head, 128-key block); dk leaves it rotated), and the dq finish.  Replaces autograd through bmm / softmax / bmm and
 * XPOS.forward of torchscale MultiheadAttention. */
int kx_attn_fwd_lse(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out, int batch,
                    int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out, kx_stream_t stream);
int kx_attn_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                const void* d_out, long long ld_dout, const float* lse, void* dq, void* dk, void* dv, long long ld_dqkv,
                float* dq_accum, float* delta, const float* xq_cos, const float* xq_sin, const float* xk_cos,
                const float* xk_sin, int batch, int heads, int seq_len, int causal, float scale, kx_stream_t stream);

/* The same pair with attention dropout (torchscale MultiheadAttention: `attn_probs = dropout(attn_weights)`; the reference
 * trains with attention_dropout = 0.1, kosmosx/model.py:177).  The keep bits are drawn AHEAD of the flash kernels — whose
 * softmax warps are their bottleneck — by kx_attn_dropout_masks, a pure function of (drop_seed, drop_site, batch*heads + head,
 * query, key): Philox4x32-7 words bit-sliced into Bernoulli(keep) bits with keep = round((1 - p) * 4096) / 4096, 1 bit per
 * score, written in the two layouts the kernels' thread mappings want (nb = ceil(seq_len / 128); causal: only tiles
 * kb <= qb are written or read; kx_attn_dropout_mask_words() uint32 words each):
 *   row_mask[(((bh * nb + qb) * nb + kb) * 128 + r) * 4 + c]       query 128 qb + r keeps key 128 kb + 32 c + k at bit
 *                                                                  (k/2)%8 + 8 (k%2) + 16 (k/16)  (even / odd keys of each
 *                                                                  16-key half in separate bytes: the forward kernel masks
 *                                                                  bf16 pairs with one shift + one byte-permute per pair)
 *   key_mask[((((bh * nb + qb) * nb + kb) * 4 + g) * 128 + r]      bit i: query 128 qb + 32 g + i keeps key 128 kb + r
 * Forward: dropped probabilities are zeroed in the P operand of P.V, the row normaliser sums ALL probabilities, the
 * output is scaled by 1 / keep.  Backward: dV += (M o P / keep)^T dO, dS = P o (M o dP / keep - delta). */
size_t kx_attn_dropout_mask_words(int batch, int heads, int seq_len);
int kx_attn_dropout_masks(float drop_p, unsigned int drop_site, unsigned long long drop_seed, int batch, int heads, int seq_len,
                          int causal, unsigned int* row_mask, unsigned int* key_mask, kx_stream_t stream);
int kx_attn_fwd_dropout(const void* q, const void* k, const void* v, long long ld_qkv, void* out, long long ld_out, int batch,
                        int heads, int seq_len, int causal, float scale, float* stats_out, float* lse_out, float drop_p,
                        const unsigned int* row_mask, kx_stream_t stream);
int kx_attn_bwd_dropout(const void* q, const void* k, const void* v, long long ld_qkv, const void* out, long long ld_out,
                        const void* d_out, long long ld_dout, const float* lse, void* dq, void* dk, void* dv, long long ld_dqkv,
                        float* dq_accum, float* delta, const float* xq_cos, const float* xq_sin, const float* xk_cos,
                        const float* xk_sin, int batch, int heads, int seq_len, int causal, float scale, float drop_p,
                        const unsigned int* key_mask, kx_stream_t stream);

/* Profiling aid for kx_attn_bwd (causal): with a device buffer of 2*32*16 + 4 * (number of CTAs) int64 installed, one CTA of
 * every launch (named by word 1023 of the buffer; negative = CTA 0) records clock64 stamps [role: compute thread 0, MMA thread]
 * [iteration][point] and every CTA b leaves {SM id, globaltimer ns at entry, at its first ready score tile, at exit} at word
 * 1024 + 4 b; NULL = off. */
int kx_attn_bwd_set_trace(long long* device_buffer);

/* out = LayerNorm(act(x)) * gamma + beta in one pass, bf16 in / bf16 out (ffn_layernorm(gelu(fc1 x)), SURVEY A.4:
 * training keeps the pre-activation u, so GELU and the LayerNorm share one read of it).  n % 8 == 0, n <= 8192. */
int kx_act_layernorm_fwd(const void* x_bf16, long long ld_x, int act, const float* gamma, const float* beta, float eps,
                         void* out_bf16, long long ld_out, int rows, int n, kx_stream_t stream);

/* nn.LayerNorm backward for y = LN(act(x)) * gamma + beta, given dy (bf16):
 *   dx = act'(x) * rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; d_gamma (+)= sum dy*xhat; d_beta (+)= sum dy.
 * dx_is_f32 = 1: dx is the fp32 residual-stream gradient, dx = dres + (LN gradient) (dres NULL = none; may alias dx),
 * dxb (NULL or bf16) receives a copy and d_colsum (NULL or fp32 [n]) the column sums of dx — the bias gradient of the
 * Linear whose output was added to the stream there.  dx_is_f32 = 0: dx is bf16 (dres/dxb must be NULL), d_colsum sums
 * the stored bf16 values.  partials: fp32 scratch [3][kx_ln_bwd_partials(rows)][n].  accumulate = 1 adds into d_*.
 * pre_add (NULL or fp32 [n], n <= 2048, act none): the LayerNorm input is x + pre_add (the perceiver's
 * `x + media_pos_emb[i]`; d_colsum is then the gradient of that table row when dres is NULL).
 * The GELU form on wide rows (2048 < n <= 8192: the decoder's ffn_layernorm) runs its own kernel, two rows in flight per CTA. */
int kx_ln_bwd_partials(int rows);
int kx_layernorm_bwd(const void* x, int x_is_bf16, long long ld_x, const float* pre_add, int act, const void* dy_bf16, long long ld_dy,
                     const float* gamma, float eps, const float* dres, long long ld_dres, void* dx, int dx_is_f32,
                     long long ld_dx, void* dxb_bf16, long long ld_dxb, float* partials, int n_partials, float* d_gamma,
                     float* d_beta, float* d_colsum, int accumulate, int rows, int n, float drop_p, unsigned int drop_site,
                     unsigned long long drop_seed, kx_stream_t stream);
/* (drop_p > 0, fp32 residual form only: the Linear output added to the stream at this point went through dropout in the
 * forward — kx_gemm_args.drop_* with the same (seed, site) — so dxb and d_colsum receive the masked, rescaled gradient while
 * dx, the residual-stream gradient, stays whole.) */

/* In-place dropout of an fp32 [rows, ld] matrix with the same mask function as kx_gemm_args.drop_*: torchscale's
 * `x = dropout(x)` closing Decoder.forward_embedding (the call at kosmosx/model.py:242-244), and — applied to the gradient
 * with the same (seed, site) — its backward.  cols % 8 == 0. */
int kx_dropout_f32(float* x, long long ld, int rows, int cols, float drop_p, unsigned int drop_site,
                   unsigned long long drop_seed, kx_stream_t stream);

/* out[n] += column sums of a bf16 matrix (bias gradients: d(bias) = sum over rows of dY). */
int kx_colsum_bf16(const void* x_bf16, long long ld, int rows, int n, float* out, kx_stream_t stream);

/* Transpose of the xPos rotation (kx_gemm_args KX_EPI_QKV_XPOS) applied in place to the q and k column blocks
 * [0, 2*d_model) of a bf16 [rows, ld] gradient matrix; tables from kx_xpos_tables. */
int kx_xpos_bwd(void* dqkv_bf16, long long ld, int rows, int d_model, int seq_len, const float* q_cos, const float* q_sin,
                const float* k_cos, const float* k_sin, kx_stream_t stream);

/* Next-token targets of the spliced sequence (same splice description as kx_embed_splice_pos), one int64 per row,
 * -100 = no loss (torch's ignore_index), and *count += rows with a target (device fp32, may be NULL).
 *   KX_LOSS_REFERENCE   the reference's intended loss, experimental/model/allModalities/notes.txt:566-574:
 *       `outputs = cat([outputs[:, :1], outputs[:, 67:]])`, `loss(outputs[:, :-1], labels[:, 1:])` with labels = the text
 *       WITHOUT the `<image>` `</image>` markers (model.py:70-77).  Generalised to any image position p (features in
 *       front of text token p): text tokens p-1 and p are the markers; marker rows and feature rows carry no loss and
 *       markers are never targets; every other text row predicts the next non-marker text token.  For the reference
 *       layout (p = 2) that is exactly rows 0 and 67.., row 0 predicting text token 3.
 *   KX_LOSS_NEXT_TOKEN  plain next-token rule: the row of text token i predicts text token i+1; feature rows, the last
 *       token and the token directly in front of an image carry no loss.
 * ignore_token >= 0: targets equal to it (the tokenizer's <pad>) are dropped as well; -1 = keep everything (the
 * reference's loop does no pad masking). */
enum { KX_LOSS_REFERENCE = 0, KX_LOSS_NEXT_TOKEN = 1 };
int kx_loss_targets(const long long* tokens, int batch, int t_text, const int* host_img_rows, int img_count, int n_img,
                    int rule, long long ignore_token, long long* targets, float* count, kx_stream_t stream);

/* Softmax cross-entropy (torch.nn.functional.cross_entropy, reduction = mean over the rows with a target):
 * loss_acc[0] += sum of row losses, loss_acc[1] += rows counted.  targets: int64 [rows], negative = ignored.
 * dlogits (NULL = loss only): bf16 [rows, ld_dlogits], (softmax - onehot) / max(*count, 1), zeros for ignored rows and
 * for the pad columns [vocab, ld_dlogits), so the matrix can feed kx_gemm_bf16 directly.  count: DEVICE fp32 scalar
 * (kx_loss_targets), read by the kernel — the normalisation needs no host value.  Targets >= vocab set *err_flag. */
int kx_ce_fwd_bwd(const float* logits, long long ld_logits, const long long* targets, int rows, int vocab,
                  const float* count, void* dlogits_bf16, long long ld_dlogits, float* loss_acc, int* err_flag,
                  kx_stream_t stream);

/* Backward of kx_embed_splice_pos: d_embed[token] += dx0[row] for text rows (not for padding_idx), d_pos[t + 2] +=
 * dx0[row] for every row, and with alias_positions also d_pos[i + 2] += dx0[row] for the row of text token i.
 * NULL tables are skipped. */
int kx_embed_bwd(const float* dx0, const long long* tokens, int batch, int t_text, const int* host_img_rows, int img_count,
                 int n_img, int dim, int vocab, int padding_idx, int alias_positions, float* d_embed, float* d_pos,
                 kx_stream_t stream);

/* Perceiver resampler pieces of the backward pass (flamingo_pytorch PerceiverResampler is trainable in the reference,
 * model.py:196-203): the cross-attention backward (dq like q; dkv like kv, k and v blocks), the GELU of its
 * feed-forward (which has no LayerNorm behind it), a row gather dst[r] (+)= src[(r/grp_rows)*grp_stride + grp_off +
 * r%grp_rows] (fp32 or bf16 -> bf16: image rows of the decoder-input gradient; media / latent rows of the
 * [media | latents] gradient) and a sum over rows in fp32 (gradient of the broadcast latents). */
int kx_perceiver_xattn_bwd(const void* q, long long ld_q, const void* kv, long long ld_kv, int v_col_off, const void* out,
                           long long ld_out, const void* d_out, long long ld_dout, void* dq, long long ld_dq, void* dkv,
                           long long ld_dkv, int batch, int heads, int n_q, int n_kv, float scale, kx_stream_t stream);
int kx_gelu_fwd(const void* u_bf16, void* out_bf16, long long n, kx_stream_t stream);
int kx_gelu_bwd(const void* u_bf16, const void* dmid_bf16, void* du_bf16, long long n, kx_stream_t stream);
/* The same pair with the activation named (KX_ACT_GELU = the two above; KX_ACT_QUICK_GELU = CLIP's x*sigmoid(1.702x),
 * [HF] activations.py QuickGELUActivation): the MLP of the last ViT layer when it is fine-tuned (notes.txt:537-538). */
int kx_act_fwd(const void* u_bf16, void* out_bf16, long long n, int act, kx_stream_t stream);
int kx_act_bwd(const void* u_bf16, const void* dmid_bf16, void* du_bf16, long long n, int act, kx_stream_t stream);
int kx_gather_rows(const void* src, int src_is_f32, long long ld_src, void* dst_bf16, long long ld_dst, int rows, int n,
                   int grp_rows, int grp_stride, int grp_off, int accumulate, kx_stream_t stream);
int kx_sum_rows_f32(const float* src, long long ld, int rows, long long n, float* out, int accumulate, kx_stream_t stream);

/* Gradient clipping (clip_grad_norm_, train.py:652-653) without a host sync: *out += sum g^2; then
 * scale = pre_scale * min(1, max_norm / (pre_scale * sqrt(sumsq) + 1e-6)), norm_out = pre_scale * sqrt(sumsq)
 * (pre_scale = 1 / world size turns all-reduced gradient sums into means; max_norm <= 0 = no clipping). */
#define KX_SUMSQ_SCRATCH 2048   /* floats of scratch kx_sumsq needs (per-block partials, folded in index order: the norm is
                                * bit-reproducible, so replicas holding identical gradients take identical steps) */
int kx_sumsq(const float* g, long long n, float* out, float* scratch, kx_stream_t stream);
int kx_clip_scale(const float* sumsq, float max_norm, float pre_scale, float* scale_out, float* norm_out, kx_stream_t stream);

/* Fused optimizers over flat fp32 buffers (master weights, gradients, moments); *grad_scale (device, may be NULL)
 * multiplies every gradient; w_bf16 (may be NULL) receives the bf16 tensor-core copy of the updated weights.
 * kx_adamw_step = torch.optim.AdamW (step >= 1 for the bias corrections); kx_lion_step = lion_pytorch.Lion. */
int kx_adamw_step(float* p, const float* g, float* m, float* v, void* w_bf16, long long n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, const float* grad_scale, kx_stream_t stream);
int kx_lion_step(float* p, const float* g, float* m, void* w_bf16, long long n, float lr, float beta1, float beta2,
                 float weight_decay, const float* grad_scale, kx_stream_t stream);

/* ==================================================================================== *
 * Incremental decoding: torchscale's `incremental_state` protocol of Decoder.forward /
 * MultiheadAttention.forward (SURVEY A.4, A.5 [recall]; reached through the decoder built at model.py:186-191).
 * The prompt pass is the ordinary forward plus kx_kv_cache_store per layer; every later step processes ONE new token
 * per sequence.  The cache holds, per layer, the xPos-ROTATED keys and the values as bf16 [batch, t_max, d_model]
 * (stored head-major, [batch, heads, t_max, 64], so one (batch, head) is one contiguous stream).  torchscale caches un-rotated keys and re-rotates the whole key sequence with a
 * re-centred scale every step; q.k depends on the position DIFFERENCE only (A.5), so rotating each key once with the
 * prompt's centre gives the same products.  The current position (= number of cached tokens) lives in DEVICE memory
 * (`pos`), so one captured CUDA graph serves every step with no host round trip.
 * ==================================================================================== */
enum { KX_DEC_PLAIN = 0, KX_DEC_RESIDUAL = 1, KX_DEC_QKV = 2 };
#define KX_DECODE_MAX_BATCH 32

/* y[batch, N] = epilogue(a[batch, K] . W[N, K]^T) for batch <= KX_DECODE_MAX_BATCH rows: the weight-streaming
 * (HBM-bound) form of kx_gemm_bf16.  ln_c != NULL folds the LayerNorm in front of the Linear exactly as
 * kx_gemm_args.ln_part does (W carries gamma, bias = W.beta + b, ln_c = row sums of the bf16 W), with mean / rstd taken
 * over the K bf16 values of each row of `a` inside the kernel.
 *   KX_DEC_PLAIN     out (bf16 or fp32) = act(y)                      fc1 (+GELU), output_projection
 *   KX_DEC_RESIDUAL  x (fp32, in place) += y; xb = bf16(x)            out_proj, fc2 (+ residual)
 *   KX_DEC_QKV       N = 3*d_model: q (rotated with xq tables at *pos) -> q_out [batch, ld_q]; k (rotated with the xk
 *                    tables) and v -> row *pos of k_cache / v_cache [batch, t_max, d_model]
 * Replaces, for one-token steps: q/k/v/out_proj, fc1, fc2 of torchscale MultiheadAttention / FeedForwardNetwork together
 * with self_attn_layer_norm / inner_attn_ln / final_layer_norm / ffn_layernorm, XPOS.forward(offset = src_len - 1), the
 * `prev_key` / `prev_value` torch.cat, decoder.layer_norm + output_projection. */
typedef struct kx_decode_linear_args {
    int mode;                     /* KX_DEC_* */
    int act;                      /* KX_ACT_* (KX_DEC_PLAIN) */
    const float* bias;            /* fp32 [N] or NULL */
    const float* ln_c;            /* fp32 [N] or NULL = no LayerNorm fold */
    float ln_eps;
    void* out; long long ld_out; int out_f32;                  /* KX_DEC_PLAIN */
    unsigned long long* argmax_keys; /* KX_DEC_PLAIN, may be NULL: [batch] order-preserving (value, index) keys, atomicMax-reduced
                                      * over N by this launch (the greedy choice fused into output_projection); zero before use */
    float* x; long long ld_x; void* xb; long long ld_xb;       /* KX_DEC_RESIDUAL */
    void* q_out; long long ld_q; void* k_cache; void* v_cache; /* KX_DEC_QKV */
    int t_max, d_model;
    const int* pos;               /* device: position of the new token = tokens already cached */
    const float *xq_cos, *xq_sin, *xk_cos, *xk_sin;            /* kx_xpos_tables with >= t_max rows */
} kx_decode_linear_args;

int kx_decode_linear(const void* a_bf16, long long lda, int batch, const void* w_bf16, long long ldw, int N, int K,
                     const kx_decode_linear_args* args, kx_stream_t stream);

/* softmax(q . K^T * scale) . V of the new token against keys 0..*pos of the cache (flash-decoding: one CTA per
 * (128-key chunk, head, batch), the last CTA of a (batch, head) merges the chunk partials).  q bf16 [batch, ld_q],
 * out bf16 [batch, ld_out], head_dim 64.  scratch: kx_decode_attn_scratch_bytes(); counters: int [batch*heads],
 * zeroed once by the caller (the kernel leaves them zero).  Replaces the bmm / softmax(fp32) / bmm of torchscale
 * MultiheadAttention when incremental_state holds prev_key / prev_value (no mask: SURVEY A.4). */
size_t kx_decode_attn_scratch_bytes(int batch, int heads, int t_max);
int kx_decode_attn(const void* q_bf16, long long ld_q, const void* k_cache, const void* v_cache, int t_max, int batch,
                   int heads, const int* pos, float scale, float* scratch, int* counters, void* out_bf16, long long ld_out,
                   kx_stream_t stream);

/* Prompt pass: copy the k and v column blocks ([d_model, 3*d_model) of a layer's rotated q|k|v matrix, rows b*seq_len+t)
 * into rows 0..seq_len-1 of the cache (`incremental_state[idx]["prev_key"/"prev_value"] = k, v` of the first step). */
int kx_kv_cache_store(const void* qkv_bf16, long long ld_qkv, int batch, int seq_len, int d_model, void* k_cache,
                      void* v_cache, int t_max, kx_stream_t stream);

/* x[b] = embed_table[tokens[b]] + pos_table[*pos + 2] (fp32) and its bf16 copy: Decoder.forward_embedding for
 * `tokens[:, -1:]` with the last position.  text_index_off >= 0 additionally adds
 * pos_table[*pos + 2 - text_index_off]: the new token continues a sequence embedded with alias_positions
 * (kx_embed_splice_pos), whose text index trails the spliced row by the feature rows in front of it, so that a
 * generation step equals Kosmos.forward over the grown text; -1 = single add (torchscale's own incremental step).
 * err_flag bit 0: token id out of range; bit 1: position table exhausted. */
int kx_decode_embed(const long long* tokens, int batch, const float* embed_table, int vocab, const float* pos_table,
                    int pos_rows, const int* pos, int text_index_off, int dim, float* x, void* xb_bf16, int* err_flag,
                    kx_stream_t stream);

/* Greedy choice on the device: tokens_out[b] = argmax_v logits[b, v] (lowest index on ties) or forced[b, *step] when
 * `forced` (int64 [batch, history_ld]) is given; history[b, *step] = the choice; then *step += 1 and *pos += 1 (pos
 * may be NULL).  counter: one int, zeroed once by the caller.  argmax_keys (may be NULL): the keys reduced by the
 * output_projection launch (kx_decode_linear_args.argmax_keys) — the logits are then not re-read and the keys are
 * reset to zero for the next step. */
int kx_argmax_advance(const float* logits, long long ld, int batch, int vocab, const long long* forced,
                      long long* tokens_out, long long* history, int history_ld, int* pos, int* step, int* counter,
                      unsigned long long* argmax_keys, kx_stream_t stream);

/* One decoding step as ONE persistent cooperative kernel (batch <= 8): embedding, every layer's q|k|v / attention /
 * out_proj / fc1 / fc2, decoder.layer_norm + output_projection and the greedy choice, separated by grid barriers, with
 * the next phase's first weight loads issued before each barrier wait.  Same arithmetic as the kx_decode_* calls above
 * (which remain the path for 8 < batch <= 32 and for per-kernel profiling).
 *   kx_decode_plan_build  one-time setup per generation: flattens the arguments into `device_plan`
 *                         (kx_decode_plan_bytes(layers) bytes).  Copies from a temporary host buffer, so it
 *                         SYNCHRONISES `stream` — the only call of this library that does.
 *   kx_decode_step        one launch = one new token per sequence; tokens[b] is consumed, the next choice is written
 *                         back to tokens[b] and history[b, *step]; *pos and *step advance.
 * Per-layer arrays are HOST arrays of `layers` DEVICE pointers.  barrier: 288 x uint64 (heads <= 256), zeroed once and then owned by
 * the kernel;
 * err_flag bit 2 = a barrier timed out (logic error, results invalid). */
typedef struct kx_decode_step_args {
    int batch, layers, d_model, ffn, heads, vocab, t_max, pos_rows;
    int text_index_off;           /* as kx_decode_embed (-1 = single positional add) */
    float eps, scale;
    const void* const* w_qkv; const float* const* c_qkv; const float* const* d_qkv;      /* [layers] */
    const void* const* w_o;   const float* const* c_o;   const float* const* d_o;
    const void* const* w_fc1; const float* const* c_fc1; const float* const* d_fc1;
    const void* const* w_fc2; const float* const* c_fc2; const float* const* d_fc2;
    void* const* k_cache; void* const* v_cache;                                          /* [layers], head-major caches */
    const void* w_out; const float* c_out; const float* d_out;                           /* d_out may be NULL */
    const float* embed_table; const float* pos_table;
    const float *xq_cos, *xq_sin, *xk_cos, *xk_sin;
    long long* tokens; float* x; void* xb; void* q; void* att; void* mid;                /* per-step activations */
    float* logits; long long ld_logits;
    unsigned long long* argmax_keys; int* pos; int* step; int* err_flag;
    const long long* forced; long long* history; int history_ld;                         /* forced / history may be NULL */
    unsigned long long* barrier;
    long long* trace;   /* optional profiling aid: int64 [2 * phases + 17]; CTA 0 stamps %globaltimer when its own work of a
                         * phase is done and when it leaves the barrier behind it (phases = 1 + 5*layers + 2); entry
                         * [2*phases] selects one Linear phase (-1 = none) whose item gets 16 finer stamps after it */
} kx_decode_step_args;

size_t kx_decode_plan_bytes(int layers);
int kx_decode_step_ctas(void);   /* grid size of the cooperative kernel on the current device (CTAs per SM x SMs), < 0 on error */
int kx_decode_plan_build(const kx_decode_step_args* args, void* device_plan, kx_stream_t stream);
int kx_decode_step(const void* device_plan, kx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KOSMOSX_B200_H */
