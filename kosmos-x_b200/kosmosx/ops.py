"""Thin Python wrappers over the C ABI: tensors in, raw pointers out.  No arithmetic here."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from ._abi import GemmArgs, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---- optional per-launch timing (bench.py's roofline leg) -------------------------------------
# When a list is installed with profile_begin(), every wrapper brackets its launch with CUDA
# events on the launching stream and appends (kind, flops, bytes, start, end).  Off by default.
_prof = None


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> list of (kind, flops, bytes, milliseconds); synchronises."""
    global _prof
    recs, _prof = _prof or [], None
    torch.cuda.synchronize()
    return [(k, fl, by, e0.elapsed_time(e1)) for k, fl, by, e0, e1 in recs]


class _Timed:
    __slots__ = ("kind", "flops", "bytes", "e0")

    def __init__(self, kind, flops=0.0, nbytes=0.0):
        self.kind, self.flops, self.bytes = kind, flops, nbytes

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _prof is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.kind, self.flops, self.bytes, self.e0, e1))
        return False


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: kosmosx has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias=None, res=None, act=_abi.KX_ACT_NONE,
         grp=None, add_tab=None, add_off=0, xpos=None, seq_len=0, cta_group=0, block_n=0, max_ctas=0, M=None,
         epi_mode=0, ln=None, stats_out=None, out2=None, a_trans=False, b_trans=False, drop=None):
    """out = epilogue(a[M,K] @ w[N,K]^T).  a, w bf16; out bf16 or fp32 (2-D views, row pitch = stride(0)).
    a_trans / b_trans: the operand is given as [K, M] / [K, N] (backward GEMMs: dgrad = gemm(dY, W, b_trans=True),
    wgrad = gemm(dY, X, a_trans=True, b_trans=True)).

    ln = (partials fp32 [tiles, M, 2], c fp32 [N], cols, eps): LayerNorm of the rows of `a` folded into the
    epilogue (w must carry gamma, bias must be W.beta + b).  stats_out fp32 [ceil(N/128), M, 2] and out2
    (bf16 copy of an fp32 out) make this GEMM the producer of the next fold.
    drop = (p, site, seed): training dropout on the Linear's output, before the residual add (kx_gemm_args.drop_*)."""
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    g = GemmArgs()
    g.M = (a.shape[1] if a_trans else a.shape[0]) if M is None else M
    g.K = a.shape[0] if a_trans else a.shape[1]
    g.N = w.shape[1] if b_trans else w.shape[0]
    if (w.shape[0] if b_trans else w.shape[1]) != g.K:
        raise ValueError(f"gemm: K mismatch {tuple(a.shape)} vs {tuple(w.shape)} (a_trans={a_trans}, b_trans={b_trans})")
    g.a_trans, g.b_trans = int(a_trans), int(b_trans)
    g.bias = _ptr(bias)
    g.res = _ptr(res)
    g.ld_res = res.stride(0) if res is not None else 0
    g.out = out.data_ptr()
    g.ld_out = out.stride(0)
    g.out_f32 = 1 if out.dtype == torch.float32 else 0
    if out.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("gemm: out must be fp32 or bf16")
    g.act = act
    if grp is not None:
        g.grp_rows, g.grp_stride, g.grp_off = grp
    g.add_tab = _ptr(add_tab)
    g.add_off = add_off
    g.ld_add = add_tab.stride(0) if add_tab is not None else 0
    if xpos is not None:
        g.epi = _abi.KX_EPI_QKV_XPOS
        g.xq_cos, g.xq_sin, g.xk_cos, g.xk_sin = (t.data_ptr() for t in xpos)
        g.seq_len = seq_len
        g.d_model = g.N // 3
    g.cta_group, g.block_n, g.max_ctas, g.epi_mode = cta_group, block_n, max_ctas, epi_mode
    if ln is not None:
        part, c, cols, eps = ln
        _req(part, torch.float32, "ln partials")
        if part.ndim != 3 or part.shape[1] != g.M or part.shape[2] != 2 or not part.is_contiguous():
            raise ValueError(f"gemm: ln partials must be contiguous [tiles, M={g.M}, 2], got {tuple(part.shape)}")
        if c.numel() != g.N:
            raise ValueError("gemm: ln c must have N entries")
        g.ln_part, g.ln_c, g.ln_tiles, g.ln_cols, g.ln_eps = part.data_ptr(), c.data_ptr(), part.shape[0], cols, eps
    if stats_out is not None:
        _req(stats_out, torch.float32, "stats_out")
        # one partial per half tile: [ceil(N/128), M, 2] selects 256-wide tiles, [ceil(N/64), M, 2] 128-wide ones
        if not stats_out.is_contiguous() or stats_out.ndim != 3 or tuple(stats_out.shape[1:]) != (g.M, 2):
            raise ValueError(f"gemm: stats_out must be contiguous [tiles, {g.M}, 2]")
        if stats_out.shape[0] == (g.N + 127) // 128 and block_n in (0, 256):
            g.block_n = 256
        elif stats_out.shape[0] == (g.N + 63) // 64 and block_n in (0, 128):
            g.block_n = 128
        else:
            raise ValueError(f"gemm: stats_out has {stats_out.shape[0]} partials; expected {(g.N + 127) // 128} "
                             f"(block_n 256) or {(g.N + 63) // 64} (block_n 128)")
        g.stats_out = stats_out.data_ptr()
    if out2 is not None:
        _req(out2, torch.bfloat16, "out2")
        g.out2, g.ld_out2 = out2.data_ptr(), out2.stride(0)
    if drop is not None and drop[0] > 0:
        g.drop_p, g.drop_site, g.drop_seed = float(drop[0]), int(drop[1]), int(drop[2])
    with _Timed(f"gemm {g.M}x{g.N}x{g.K}" + ("+tn" if a_trans else "+nn" if b_trans else "") + ("+ln" if ln is not None else "") + ("+xpos" if xpos is not None else "")
                + ("+gelu" if act == _abi.KX_ACT_GELU else "") + ("+res" if res is not None else ""),
                2.0 * g.M * g.N * g.K,
                2.0 * (g.M + g.N) * g.K + g.M * g.N * (out.element_size() + (4 if res is not None else 0))):
        check(lib.kx_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), g, _stream()), "kx_gemm_bf16")
    return out


def attention(q, k, v, out, *, batch, heads, seq_len, causal, scale, stats_out=None, lse_out=None, drop_p=0.0, row_mask=None):
    """q, k, v: bf16 2-D views [batch*seq_len, heads*64] sharing one row pitch; out bf16 [batch*seq_len, >=heads*64].
    stats_out fp32 [heads, batch*seq_len, 2]: per-head partial (sum, sumsq) of every output row."""
    if stats_out is not None:
        _req(stats_out, torch.float32, "stats_out")
        if tuple(stats_out.shape) != (heads, batch * seq_len, 2) or not stats_out.is_contiguous():
            raise ValueError("attention: stats_out must be contiguous [heads, batch*seq_len, 2]")
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        _req(t, torch.bfloat16, n)
    if not (q.stride(0) == k.stride(0) == v.stride(0)):
        raise ValueError("attention: q, k, v must share a row pitch")
    fl = 4.0 * batch * heads * seq_len * seq_len * 64 * (0.5 if causal else 1.0)
    with _Timed("attn_causal" if causal else "attn_full", fl, 8.0 * batch * heads * seq_len * 64):
        if lse_out is not None:
            _req(lse_out, torch.float32, "lse_out")
            if lse_out.numel() != heads * batch * lse_pad(seq_len) or not lse_out.is_contiguous():
                raise ValueError("attention: lse_out must be contiguous [heads, batch, ceil(seq_len/128)*128]")
        if drop_p > 0:                                # training: attention dropout, keep bits from attn_dropout_masks()
            if lse_out is None or row_mask is None or row_mask.dtype != torch.int32 or row_mask.numel() < attn_dropout_mask_words(batch, heads, seq_len):
                raise ValueError("attention: dropout needs lse_out and the int32 row_mask of attn_dropout_masks()")
            check(lib.kx_attn_fwd_dropout(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), out.data_ptr(), out.stride(0),
                                          batch, heads, seq_len, 1 if causal else 0, float(scale), _ptr(stats_out), lse_out.data_ptr(),
                                          float(drop_p), row_mask.data_ptr(), _stream()), "kx_attn_fwd_dropout")
        elif lse_out is not None:
            check(lib.kx_attn_fwd_lse(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), out.data_ptr(), out.stride(0),
                                      batch, heads, seq_len, 1 if causal else 0, float(scale), _ptr(stats_out),
                                      lse_out.data_ptr(), _stream()), "kx_attn_fwd_lse")
        else:
            check(lib.kx_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), out.data_ptr(), out.stride(0),
                                  batch, heads, seq_len, 1 if causal else 0, float(scale), _ptr(stats_out), _stream()),
                  "kx_attn_fwd")
    return out


def lse_pad(seq_len: int) -> int:
    return (seq_len + 127) // 128 * 128


def attn_dropout_mask_words(batch: int, heads: int, seq_len: int) -> int:
    return int(lib.kx_attn_dropout_mask_words(batch, heads, seq_len))


def attn_dropout_masks(row_mask, key_mask, *, p, site, seed, batch, heads, seq_len, causal=True):
    """Draw the keep bits of one attention-dropout site into both layouts (kx_attn_dropout_masks)."""
    for n, t in (("row_mask", row_mask), ("key_mask", key_mask)):
        if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous() or t.numel() < attn_dropout_mask_words(batch, heads, seq_len):
            raise ValueError(f"attn_dropout_masks: {n} must be a contiguous int32 CUDA tensor of attn_dropout_mask_words() entries")
    with _Timed("attn_dropout_masks", 0.0, 8.0 * batch * heads * seq_len * seq_len / 8 / 2):
        check(lib.kx_attn_dropout_masks(float(p), int(site), int(seed), batch, heads, seq_len, 1 if causal else 0,
                                        row_mask.data_ptr(), key_mask.data_ptr(), _stream()), "kx_attn_dropout_masks")


def unpack_attn_row_mask(mask: torch.Tensor, batch: int, heads: int, seq_len: int) -> torch.Tensor:
    """row_mask of attn_dropout_masks -> bool (batch, heads, seq_len, seq_len) [q, k] (test helper)."""
    nb = (seq_len + 127) // 128
    w = mask[:batch * heads * nb * nb * 512].view(batch, heads, nb, nb, 128, 4).to(torch.int64) & 0xffffffff     # [b,h,qb,kb,r,c]
    # key k = 2j + e of a word sits at bit (j % 8) + 8 e + 16 (j / 8): the pair layout of attention.cu
    k = torch.arange(32, device=mask.device)
    pos = (k // 2) % 8 + 8 * (k % 2) + 16 * (k // 16)
    bits = (w.unsqueeze(-1) >> pos) & 1                                                                          # [..., r, c, key]
    keep = bits.permute(0, 1, 2, 4, 3, 5, 6).reshape(batch, heads, nb * 128, nb * 128)                             # q = qb,r ; k = kb,c,key
    return keep[:, :, :seq_len, :seq_len].bool()


def unpack_attn_dropout_mask(mask: torch.Tensor, batch: int, heads: int, seq_len: int) -> torch.Tensor:
    """key_mask of attn_dropout_masks -> bool (batch, heads, seq_len, seq_len) [q, k] (test / inspection helper;
    the kernels read the packed words).  Entries of tiles the causal forward never visits are undefined."""
    nb = (seq_len + 127) // 128
    w = mask[:batch * heads * nb * nb * 512].view(batch, heads, nb, nb, 4, 128).to(torch.int64) & 0xffffffff      # [b,h,qb,kb,g,r]
    bits = (w.unsqueeze(-1) >> torch.arange(32, device=mask.device)) & 1                                         # [..., g, r, i]
    keep = bits.permute(0, 1, 2, 4, 6, 3, 5).reshape(batch, heads, nb * 128, nb * 128)                             # q = qb,g,i ; k = kb,r
    return keep[:, :, :seq_len, :seq_len].bool()


def dropout_f32(x, *, p, site, seed):
    """In-place dropout of an fp32 2-D matrix with the library's mask function (kx_dropout_f32)."""
    _req(x, torch.float32, "x")
    check(lib.kx_dropout_f32(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], float(p), int(site), int(seed), _stream()),
          "kx_dropout_f32")
    return x


# ---- training step ---------------------------------------------------------------------------
def attention_bwd(q, k, v, out, d_out, lse, dq, dk, dv, dq_accum, delta, *, batch, heads, seq_len, causal, scale, xpos=None,
                  drop_p=0.0, drop_mask=None):
    """dq, dk, dv (bf16 column blocks sharing one pitch) from d_out; xpos = the four kx_xpos_tables undoes the rotation."""
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out), ("d_out", d_out), ("dq", dq), ("dk", dk), ("dv", dv)):
        _req(t, torch.bfloat16, n)
    _req(lse, torch.float32, "lse"); _req(dq_accum, torch.float32, "dq_accum"); _req(delta, torch.float32, "delta")
    if not (q.stride(0) == k.stride(0) == v.stride(0)) or not (dq.stride(0) == dk.stride(0) == dv.stride(0)):
        raise ValueError("attention_bwd: q/k/v and dq/dk/dv must each share a row pitch")
    if dq_accum.numel() != batch * seq_len * heads * 64 or delta.numel() != 2 * lse.numel():
        raise ValueError("attention_bwd: scratch buffers have the wrong size")
    tabs = [None] * 4 if xpos is None else [t.data_ptr() for t in xpos]
    fl = 10.0 * batch * heads * seq_len * seq_len * 64 * (0.5 if causal else 1.0)
    with _Timed("attn_bwd", fl, 0.0):
        if drop_mask is not None and drop_p > 0:
            check(lib.kx_attn_bwd_dropout(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), out.data_ptr(), out.stride(0),
                                          d_out.data_ptr(), d_out.stride(0), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                          dv.data_ptr(), dq.stride(0), dq_accum.data_ptr(), delta.data_ptr(), *tabs, batch,
                                          heads, seq_len, 1 if causal else 0, float(scale), float(drop_p),
                                          drop_mask.data_ptr(), _stream()), "kx_attn_bwd_dropout")
        else:
            check(lib.kx_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), out.data_ptr(), out.stride(0),
                                  d_out.data_ptr(), d_out.stride(0), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                  dq.stride(0), dq_accum.data_ptr(), delta.data_ptr(), *tabs, batch, heads, seq_len,
                                  1 if causal else 0, float(scale), _stream()), "kx_attn_bwd")


def act_layernorm(x, gamma, beta, out, *, act=_abi.KX_ACT_GELU, eps=1e-5):
    _req(x, torch.bfloat16, "x"); _req(out, torch.bfloat16, "out")
    with _Timed("act_layernorm", 0.0, 4.0 * x.shape[0] * x.shape[1]):
        check(lib.kx_act_layernorm_fwd(x.data_ptr(), x.stride(0), act, gamma.data_ptr(), beta.data_ptr(), float(eps),
                                       out.data_ptr(), out.stride(0), x.shape[0], x.shape[1], _stream()), "kx_act_layernorm_fwd")
    return out


def ln_bwd_partials(rows: int) -> int:
    n = int(lib.kx_ln_bwd_partials(rows))
    if n <= 0:
        check(n, "kx_ln_bwd_partials")
    return n


def layernorm_bwd(x, dy, gamma, dx, d_gamma, d_beta, partials, *, act=_abi.KX_ACT_NONE, eps=1e-5, dres=None, dxb=None,
                  d_colsum=None, accumulate=False, pre_add=None, drop=None):
    """LayerNorm (+GELU) backward; see kx_layernorm_bwd.  partials: fp32 [3, ln_bwd_partials(rows), n]."""
    _req(dy, torch.bfloat16, "dy")
    rows, n = x.shape
    if partials.ndim != 3 or partials.shape[0] != 3 or partials.shape[2] != n or not partials.is_contiguous():
        raise ValueError("layernorm_bwd: partials must be contiguous fp32 [3, P, n]")
    with _Timed("layernorm_bwd", 0.0, float(rows) * n * (x.element_size() + 2 + dx.element_size())):
        check(lib.kx_layernorm_bwd(x.data_ptr(), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), _ptr(pre_add), act, dy.data_ptr(),
                                   dy.stride(0), gamma.data_ptr(), float(eps), _ptr(dres), 0 if dres is None else dres.stride(0),
                                   dx.data_ptr(), 1 if dx.dtype == torch.float32 else 0, dx.stride(0), _ptr(dxb),
                                   0 if dxb is None else dxb.stride(0), partials.data_ptr(), partials.shape[1],
                                   d_gamma.data_ptr(), d_beta.data_ptr(), _ptr(d_colsum), 1 if accumulate else 0, rows, n,
                                   *((float(drop[0]), int(drop[1]), int(drop[2])) if drop is not None and drop[0] > 0 else (0.0, 0, 0)),
                                   _stream()), "kx_layernorm_bwd")
    return dx


def perceiver_attention_bwd(q, kv, out, d_out, dq, dkv, *, batch, heads, n_q, n_kv, v_col_off, scale):
    for n, t in (("q", q), ("kv", kv), ("out", out), ("d_out", d_out), ("dq", dq), ("dkv", dkv)):
        _req(t, torch.bfloat16, n)
    with _Timed("perceiver_xattn_bwd", 10.0 * batch * heads * n_q * n_kv * 64):
        check(lib.kx_perceiver_xattn_bwd(q.data_ptr(), q.stride(0), kv.data_ptr(), kv.stride(0), v_col_off, out.data_ptr(),
                                         out.stride(0), d_out.data_ptr(), d_out.stride(0), dq.data_ptr(), dq.stride(0),
                                         dkv.data_ptr(), dkv.stride(0), batch, heads, n_q, n_kv, float(scale), _stream()),
              "kx_perceiver_xattn_bwd")


def gelu_fwd(u, out, act=_abi.KX_ACT_GELU):
    """out = act(u), bf16 (act: KX_ACT_GELU or KX_ACT_QUICK_GELU)."""
    _req(u, torch.bfloat16, "u"); _req(out, torch.bfloat16, "out")
    check(lib.kx_act_fwd(u.data_ptr(), out.data_ptr(), u.numel(), int(act), _stream()), "kx_act_fwd")
    return out


def gelu_bwd(u, dmid, du, act=_abi.KX_ACT_GELU):
    for n, t in (("u", u), ("dmid", dmid), ("du", du)):
        _req(t, torch.bfloat16, n)
    check(lib.kx_act_bwd(u.data_ptr(), dmid.data_ptr(), du.data_ptr(), u.numel(), int(act), _stream()), "kx_act_bwd")
    return du


def gather_rows(src, dst, *, grp=None, accumulate=False):
    """dst[r] (+)= src[(r // grp_rows) * grp_stride + grp_off + r % grp_rows]; src fp32 or bf16, dst bf16."""
    _req(dst, torch.bfloat16, "dst")
    g = grp or (0, 0, 0)
    check(lib.kx_gather_rows(src.data_ptr(), 1 if src.dtype == torch.float32 else 0, src.stride(0), dst.data_ptr(), dst.stride(0),
                             dst.shape[0], dst.shape[1], g[0], g[1], g[2], 1 if accumulate else 0, _stream()), "kx_gather_rows")
    return dst


def sum_rows_f32(src, out, *, accumulate=False):
    _req(src, torch.float32, "src"); _req(out, torch.float32, "out")
    check(lib.kx_sum_rows_f32(src.data_ptr(), src.stride(0), src.shape[0], src.shape[1], out.data_ptr(), 1 if accumulate else 0,
                              _stream()), "kx_sum_rows_f32")
    return out


def colsum(x, out):
    _req(x, torch.bfloat16, "x"); _req(out, torch.float32, "out")
    with _Timed("colsum", 0.0, 2.0 * x.shape[0] * x.shape[1]):
        check(lib.kx_colsum_bf16(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), _stream()), "kx_colsum_bf16")
    return out


def xpos_bwd(dqkv, d_model, seq_len, tabs):
    _req(dqkv, torch.bfloat16, "dqkv")
    check(lib.kx_xpos_bwd(dqkv.data_ptr(), dqkv.stride(0), dqkv.shape[0], d_model, seq_len, *[t.data_ptr() for t in tabs],
                          _stream()), "kx_xpos_bwd")
    return dqkv


def _rows(img_rows):
    return (_abi.C.c_int * max(1, len(img_rows)))(*[int(r) for r in img_rows])


_LOSS_RULES = {"reference": _abi.KX_LOSS_REFERENCE, "next_token": _abi.KX_LOSS_NEXT_TOKEN}


def loss_targets(tokens, targets, count, *, img_rows=(), n_img=0, rule="reference", ignore_token=None):
    """targets int64 [B, T] (-100 = no loss) for the spliced sequence; count fp32 [1] += rows with a target."""
    _req(tokens, torch.int64, "tokens"); _req(targets, torch.int64, "targets")
    if rule not in _LOSS_RULES:
        raise ValueError(f"loss rule must be one of {sorted(_LOSS_RULES)}")
    B, t_text = tokens.shape
    if targets.numel() != B * (t_text + n_img * len(img_rows)) or not targets.is_contiguous():
        raise ValueError("loss_targets: targets must be contiguous [B, T]")
    check(lib.kx_loss_targets(tokens.data_ptr(), B, t_text, _rows(img_rows), len(img_rows), n_img, _LOSS_RULES[rule],
                              -1 if ignore_token is None else int(ignore_token), targets.data_ptr(), _ptr(count), _stream()),
          "kx_loss_targets")
    return targets


def ce_fwd_bwd(logits, targets, loss_acc, *, count=None, dlogits=None, err_flag=None):
    """loss_acc fp32 [2] += (sum of row losses, rows counted); dlogits bf16 [rows, ld >= vocab (multiple of 8)] or None;
    count: device fp32 scalar, the gradient is divided by max(count, 1)."""
    _req(logits, torch.float32, "logits"); _req(targets, torch.int64, "targets"); _req(loss_acc, torch.float32, "loss_acc")
    if targets.numel() != logits.shape[0] or not targets.is_contiguous():
        raise ValueError("ce_fwd_bwd: one target per logits row")
    with _Timed("ce_fwd_bwd", 0.0, 10.0 * logits.shape[0] * logits.shape[1]):
        check(lib.kx_ce_fwd_bwd(logits.data_ptr(), logits.stride(0), targets.data_ptr(), logits.shape[0], logits.shape[1],
                                _ptr(count), _ptr(dlogits), 0 if dlogits is None else dlogits.stride(0), loss_acc.data_ptr(),
                                _ptr(err_flag), _stream()), "kx_ce_fwd_bwd")
    return loss_acc


def embed_bwd(dx0, tokens, d_embed, d_pos, *, img_rows=(), n_img=0, padding_idx=1, alias_positions=False):
    _req(dx0, torch.float32, "dx0"); _req(tokens, torch.int64, "tokens")
    B, t_text = tokens.shape
    dim = dx0.shape[-1]
    vocab = d_embed.shape[0] if d_embed is not None else 0
    with _Timed("embed_bwd", 0.0, 12.0 * dx0.numel()):
        check(lib.kx_embed_bwd(dx0.data_ptr(), tokens.data_ptr(), B, t_text, _rows(img_rows), len(img_rows), n_img, dim,
                               vocab, padding_idx, 1 if alias_positions else 0, _ptr(d_embed), _ptr(d_pos), _stream()), "kx_embed_bwd")


_sumsq_scratch = {}


def sumsq(g, out):
    """out += sum g^2, bit-reproducible (per-block partials folded in a fixed order)."""
    sc = _sumsq_scratch.get(g.device)
    if sc is None:
        sc = _sumsq_scratch[g.device] = torch.empty(_abi.KX_SUMSQ_SCRATCH, dtype=torch.float32, device=g.device)
    check(lib.kx_sumsq(g.data_ptr(), g.numel(), out.data_ptr(), sc.data_ptr(), _stream()), "kx_sumsq")


def clip_scale(sumsq_t, max_norm, pre_scale, scale_out, norm_out=None):
    check(lib.kx_clip_scale(sumsq_t.data_ptr(), float(max_norm), float(pre_scale), scale_out.data_ptr(), _ptr(norm_out),
                            _stream()), "kx_clip_scale")


def adamw_step(p, g, m, v, wb, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, step=1, grad_scale=None):
    with _Timed("optimizer", 0.0, 30.0 * p.numel()):
        check(lib.kx_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(wb), p.numel(), float(lr),
                                float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
                                _ptr(grad_scale), _stream()), "kx_adamw_step")


def lion_step(p, g, m, wb, *, lr, betas=(0.9, 0.99), weight_decay=0.0, grad_scale=None):
    with _Timed("optimizer", 0.0, 22.0 * p.numel()):
        check(lib.kx_lion_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), _ptr(wb), p.numel(), float(lr), float(betas[0]),
                               float(betas[1]), float(weight_decay), _ptr(grad_scale), _stream()), "kx_lion_step")


def perceiver_attention(q, kv, out, *, batch, heads, n_q, n_kv, v_col_off, scale):
    for n, t in (("q", q), ("kv", kv), ("out", out)):
        _req(t, torch.bfloat16, n)
    with _Timed("perceiver_xattn", 4.0 * batch * heads * n_q * n_kv * 64):
        check(lib.kx_perceiver_xattn_fwd(q.data_ptr(), q.stride(0), kv.data_ptr(), kv.stride(0), v_col_off,
                                         out.data_ptr(), out.stride(0), batch, heads, n_q, n_kv, float(scale),
                                         _stream()), "kx_perceiver_xattn_fwd")
    return out


def layernorm(x, gamma, beta, out, *, eps=1e-5, pre_add=None, pre_add_group=0, grp=None, rows=None):
    """out (bf16 or fp32) = LayerNorm(x (+ pre_add)) over the last dim; x fp32 or bf16, 2-D."""
    if x.dtype not in (torch.float32, torch.bfloat16) or out.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("layernorm: x / out must be fp32 or bf16")
    _req(x, x.dtype, "x")
    _req(out, out.dtype, "out")
    g = grp or (0, 0, 0)
    r = x.shape[0] if rows is None else rows
    pa_rows = 0 if pre_add is None else (pre_add.numel() // x.shape[1])
    with _Timed("layernorm", 0.0, float(r) * x.shape[1] * (x.element_size() + out.element_size())):
        check(lib.kx_layernorm_fwd(x.data_ptr(), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), _ptr(pre_add),
                                   pre_add_group, pa_rows, gamma.data_ptr(), beta.data_ptr(), float(eps),
                                   out.data_ptr(), 1 if out.dtype == torch.float32 else 0, out.stride(0), r,
                                   x.shape[1], g[0], g[1], g[2], _stream()), "kx_layernorm_fwd")
    return out


def rowstats_cast(x, xb, stats):
    """xb = bf16(x); stats[0, m] = (sum, sumsq) of xb's row m.  x fp32 [rows, n], stats fp32 [1, rows, 2]."""
    _req(x, torch.float32, "x")
    _req(xb, torch.bfloat16, "xb")
    _req(stats, torch.float32, "stats")
    rows, n = x.shape
    if stats.numel() != rows * 2 or not stats.is_contiguous():
        raise ValueError("rowstats_cast: stats must be contiguous [1, rows, 2]")
    with _Timed("rowstats_cast", 0.0, 6.0 * rows * n):
        check(lib.kx_rowstats_cast(x.data_ptr(), x.stride(0), xb.data_ptr(), xb.stride(0), stats.data_ptr(), rows, n,
                                   _stream()), "kx_rowstats_cast")
    return xb


def embed_splice_pos(tokens, embed_table, pos_table, x0, *, img_rows=(), n_img=0, err_flag=None, alias_positions=False):
    """Text rows of the spliced sequence: gather + position.  img_rows = first spliced row of every image
    (ascending; the reference is (2,)); each image takes n_img rows, written by the image_proj GEMM.
    alias_positions: text token i at row t gets pos[i+2] + pos[t+2] (torchscale's in-place `x += positions`)."""
    _req(tokens, torch.int64, "tokens")
    B, t_text = tokens.shape
    if not tokens.is_contiguous():
        tokens = tokens.contiguous()
    rows = (_abi.C.c_int * max(1, len(img_rows)))(*[int(r) for r in img_rows])
    with _Timed("embed_splice_pos", 0.0, 12.0 * B * t_text * embed_table.shape[1]):
        check(lib.kx_embed_splice_pos(tokens.data_ptr(), B, t_text, embed_table.data_ptr(), embed_table.shape[0],
                                      _ptr(pos_table), 0 if pos_table is None else pos_table.shape[0],
                                      embed_table.shape[1], rows, len(img_rows), n_img, 1 if alias_positions else 0,
                                      x0.data_ptr(), _ptr(err_flag), _stream()), "kx_embed_splice_pos")
    return x0


def add_positions(x_in, x_out, pos_table):
    B, T, D = x_in.shape
    check(lib.kx_add_positions(x_in.data_ptr(), x_out.data_ptr(), B, T, D, pos_table.data_ptr(), pos_table.shape[0],
                               _stream()), "kx_add_positions")
    return x_out


# CLIPImageProcessor defaults of the reference's checkpoint (laion/CLIP-ViT-L-14-laion2B-s32B-b82K, model.py:36-38):
# OPENAI_CLIP_MEAN / OPENAI_CLIP_STD of HF image_utils.py
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _u8_layout(pixels, image):
    """(N,3,H,W) -> 0, (N,H,W,3) -> 1 for a contiguous uint8 tensor of model-sized images."""
    _req(pixels, torch.uint8, "pixels")
    if pixels.ndim != 4 or not pixels.is_contiguous():
        raise ValueError("uint8 pixels must be a contiguous 4-D tensor")
    if tuple(pixels.shape[1:]) == (3, image, image):
        return 0
    if tuple(pixels.shape[1:]) == (image, image, 3):
        return 1
    raise ValueError(f"uint8 pixels must be (N,3,{image},{image}) or (N,{image},{image},3), got {tuple(pixels.shape)}")


def _f3(v, what):
    if len(v) != 3:
        raise ValueError(f"{what} needs 3 channel values")
    return (C.c_float * 3)(*[float(a) for a in v])


def clip_normalize_u8(pixels, out=None, *, image, mean=CLIP_MEAN, std=CLIP_STD):
    """uint8 (N,3,H,W) / (N,H,W,3) -> fp32 pixel_values (N,3,H,W): CLIPImageProcessor's rescale + normalise
    (reference model.py:81-97) on the device, for images that already have the model's size."""
    cl = _u8_layout(pixels, image)
    if pixels.data_ptr() % 4:
        pixels = pixels.clone()
    if out is None:
        out = torch.empty(pixels.shape[0], 3, image, image, dtype=torch.float32, device=pixels.device)
    _req(out, torch.float32, "pixel_values")
    with _Timed("clip_normalize", 0.0, 5.0 * pixels.numel()):
        check(lib.kx_clip_normalize_u8(pixels.data_ptr(), cl, pixels.shape[0], image, _f3(mean, "mean"), _f3(std, "std"),
                                       out.data_ptr(), _stream()), "kx_clip_normalize_u8")
    return out


def im2col_patches_u8(pixels, patches, class_embedding, pos_table, x, *, image, patch, media=1, mean=CLIP_MEAN,
                      std=CLIP_STD):
    """im2col_patches with clip_normalize_u8 fused in front: uint8 pixels, N = sequences*media in (sequence, media) order."""
    cl = _u8_layout(pixels, image)
    if pixels.data_ptr() % 4:
        pixels = pixels.clone()                       # an offset view: the kernel reads 4-byte words
    with _Timed("im2col", 0.0, 3.0 * pixels.numel()):
        check(lib.kx_im2col_patches_u8(pixels.data_ptr(), cl, _f3(mean, "mean"), _f3(std, "std"), pixels.shape[0], media,
                                       image, patch, patches.data_ptr(), patches.shape[1], class_embedding.data_ptr(),
                                       pos_table.data_ptr(), x.data_ptr(), x.shape[-1], _stream()),
              "kx_im2col_patches_u8")
    return patches


def im2col_patches(pixels, patches, class_embedding, pos_table, x, *, image, patch, media=1):
    """pixels fp32 (N,3,H,W), N = sequences*media in (sequence, media) order; output slots are media-major."""
    _req(pixels, torch.float32, "pixels")
    if pixels.data_ptr() % 16:
        pixels = pixels.clone()                       # an offset view: the kernel reads 16-byte vectors
    with _Timed("im2col", 0.0, 6.0 * pixels.numel()):
        check(lib.kx_im2col_patches(pixels.data_ptr(), pixels.shape[0], media, image, patch, patches.data_ptr(),
                                    patches.shape[1], class_embedding.data_ptr(), pos_table.data_ptr(), x.data_ptr(),
                                    x.shape[-1], _stream()), "kx_im2col_patches")
    return patches


def xpos_tables(scale, inv_freq, T, min_pos, scale_base, device):
    tabs = torch.empty(4, T, 32, dtype=torch.float32, device=device)
    check(lib.kx_xpos_tables(scale.data_ptr(), inv_freq.data_ptr(), T, min_pos, float(scale_base), tabs[0].data_ptr(),
                             tabs[1].data_ptr(), tabs[2].data_ptr(), tabs[3].data_ptr(), _stream()), "kx_xpos_tables")
    return tabs


def cast_bf16(src: torch.Tensor, dst: torch.Tensor | None = None) -> torch.Tensor:
    _req(src, torch.float32, "src")
    src = src.contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    check(lib.kx_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "kx_cast_f32_to_bf16")
    return dst


def cast_f32(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    _req(src, torch.bfloat16, "src"); _req(dst, torch.float32, "dst")
    if src.numel() != dst.numel() or not src.is_contiguous() or not dst.is_contiguous():
        raise ValueError("cast_f32: src and dst must be contiguous and of equal size")
    check(lib.kx_cast_bf16_to_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "kx_cast_bf16_to_f32")
    return dst


def broadcast_rows(src, dst, copies):
    check(lib.kx_broadcast_rows(src.data_ptr(), dst.data_ptr(), src.numel(), copies, _stream()), "kx_broadcast_rows")
    return dst


# ---- verification precision (bf16x3) -------------------------------------------------------------
def split_bf16x3(src, dst=None, *, weights=False, n_pad=None):
    """fp32 [rows, n] -> bf16 [rows, 3*n_pad]: (hi | hi | lo) for activations, (hi | lo | hi) for weights, so that ONE
    kx_gemm_bf16 over K' = 3*n_pad accumulates Ah.Wh + Ah.Wl + Al.Wh in the same TMEM tile."""
    _req(src, torch.float32, "src")
    rows, n = src.shape
    n_pad = (n + 7) // 8 * 8 if n_pad is None else n_pad
    if dst is None:
        dst = torch.empty(rows, 3 * n_pad, dtype=torch.bfloat16, device=src.device)
    _req(dst, torch.bfloat16, "dst")
    with _Timed("split_bf16x3", 0.0, 10.0 * rows * n):
        check(lib.kx_split_bf16x3(src.data_ptr(), src.stride(0), rows, n, n_pad, dst.data_ptr(), dst.stride(0),
                                  1 if weights else 0, _stream()), "kx_split_bf16x3")
    return dst


def attention_f32(q, k, v, out, *, batch, heads, n_q, n_kv, causal, scale):
    """fp32 softmax(q.k^T*scale).v, head_dim 64; q [batch*n_q, ld], k / v [batch*n_kv, ld_kv] column blocks."""
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        _req(t, torch.float32, n)
    if k.stride(0) != v.stride(0):
        raise ValueError("attention_f32: k and v must share a row pitch")
    with _Timed("attn_f32", 4.0 * batch * heads * n_q * n_kv * 64 * (0.5 if causal else 1.0)):
        check(lib.kx_attn_f32(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), out.data_ptr(), out.stride(0),
                              batch, heads, n_q, n_kv, 1 if causal else 0, float(scale), _stream()), "kx_attn_f32")
    return out


def xpos_apply_f32(qkv, d_model, seq_len, tabs):
    _req(qkv, torch.float32, "qkv")
    check(lib.kx_xpos_apply_f32(qkv.data_ptr(), qkv.stride(0), qkv.shape[0], d_model, seq_len, *[t.data_ptr() for t in tabs],
                                _stream()), "kx_xpos_apply_f32")
    return qkv


def im2col_patches_f32(pixels, patches, class_embedding, pos_table, x, *, image, patch, media=1):
    _req(pixels, torch.float32, "pixels"); _req(patches, torch.float32, "patches")
    check(lib.kx_im2col_patches_f32(pixels.data_ptr(), pixels.shape[0], media, image, patch, patches.data_ptr(), patches.shape[1],
                                    class_embedding.data_ptr(), pos_table.data_ptr(), x.data_ptr(), x.shape[-1], _stream()),
          "kx_im2col_patches_f32")
    return patches


# ---- incremental decoding ----------------------------------------------------------------------
def decode_linear(a, w, *, bias=None, ln_c=None, eps=1e-5, act=_abi.KX_ACT_NONE, out=None, res=None, qkv=None,
                  argmax_keys=None):
    """One-token Linear for batch <= 32 rows (kx_decode_linear).  Exactly one of:
      out=tensor (bf16 / fp32 [B, >=N])            plain (+act)
      res=(x fp32 [B, N], xb bf16 [B, N])          x += y in place, xb = bf16(x)
      qkv=(q_out, k_cache, v_cache, t_max, pos, tabs)   q|k|v with xPos at *pos, k / v written into the cache."""
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    g = _abi.DecodeLinearArgs()
    g.act, g.bias, g.ln_c, g.ln_eps = act, _ptr(bias), _ptr(ln_c), eps
    N, K = w.shape
    if a.shape[1] != K:
        raise ValueError(f"decode_linear: K mismatch {tuple(a.shape)} vs {tuple(w.shape)}")
    if out is not None:
        g.mode, g.out, g.ld_out, g.out_f32 = _abi.KX_DEC_PLAIN, out.data_ptr(), out.stride(0), int(out.dtype == torch.float32)
        g.argmax_keys = _ptr(argmax_keys)
        if out.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("decode_linear: out must be fp32 or bf16")
    elif res is not None:
        x, xb = res
        _req(x, torch.float32, "x"); _req(xb, torch.bfloat16, "xb")
        g.mode, g.x, g.ld_x, g.xb, g.ld_xb = _abi.KX_DEC_RESIDUAL, x.data_ptr(), x.stride(0), xb.data_ptr(), xb.stride(0)
    elif qkv is not None:
        q_out, k_cache, v_cache, t_max, pos, tabs = qkv
        _req(q_out, torch.bfloat16, "q_out"); _req(k_cache, torch.bfloat16, "k_cache"); _req(v_cache, torch.bfloat16, "v_cache")
        if tabs.shape[1] < t_max:
            raise ValueError("decode_linear: the xPos tables must cover t_max positions")
        g.mode, g.q_out, g.ld_q = _abi.KX_DEC_QKV, q_out.data_ptr(), q_out.stride(0)
        g.k_cache, g.v_cache, g.t_max, g.d_model, g.pos = k_cache.data_ptr(), v_cache.data_ptr(), t_max, N // 3, pos.data_ptr()
        g.xq_cos, g.xq_sin, g.xk_cos, g.xk_sin = (t.data_ptr() for t in tabs)
    else:
        raise ValueError("decode_linear: give out, res or qkv")
    with _Timed(f"decode_linear {N}x{K}", 2.0 * a.shape[0] * N * K, 2.0 * N * K):
        check(lib.kx_decode_linear(a.data_ptr(), a.stride(0), a.shape[0], w.data_ptr(), w.stride(0), N, K, g, _stream()),
              "kx_decode_linear")


def decode_attn_scratch(batch, heads, t_max, device):
    n = int(lib.kx_decode_attn_scratch_bytes(batch, heads, t_max))
    return (torch.empty(max(n // 4, 1), dtype=torch.float32, device=device),
            torch.zeros(batch * heads, dtype=torch.int32, device=device))


def decode_attention(q, k_cache, v_cache, out, *, t_max, heads, pos, scale, scratch, counters):
    _req(q, torch.bfloat16, "q"); _req(out, torch.bfloat16, "out")
    B = q.shape[0]
    with _Timed("decode_attn", 0.0, 0.0):
        check(lib.kx_decode_attn(q.data_ptr(), q.stride(0), k_cache.data_ptr(), v_cache.data_ptr(), t_max, B, heads,
                                 pos.data_ptr(), float(scale), scratch.data_ptr(), counters.data_ptr(), out.data_ptr(),
                                 out.stride(0), _stream()), "kx_decode_attn")
    return out


def kv_cache_store(qkv, k_cache, v_cache, *, batch, seq_len, d_model, t_max):
    _req(qkv, torch.bfloat16, "qkv")
    with _Timed("kv_cache_store", 0.0, 8.0 * batch * seq_len * d_model):
        check(lib.kx_kv_cache_store(qkv.data_ptr(), qkv.stride(0), batch, seq_len, d_model, k_cache.data_ptr(),
                                    v_cache.data_ptr(), t_max, _stream()), "kx_kv_cache_store")


def decode_embed(tokens, embed_table, pos_table, pos, x, xb, err_flag=None, text_index_off=-1):
    _req(tokens, torch.int64, "tokens")
    check(lib.kx_decode_embed(tokens.data_ptr(), tokens.numel(), embed_table.data_ptr(), embed_table.shape[0],
                              pos_table.data_ptr(), pos_table.shape[0], pos.data_ptr(), int(text_index_off), embed_table.shape[1], x.data_ptr(),
                              xb.data_ptr(), _ptr(err_flag), _stream()), "kx_decode_embed")


def argmax_advance(logits, tokens_out, *, step, counter, pos=None, history=None, forced=None, keys=None):
    _req(logits, torch.float32, "logits")
    hist_ld = history.shape[1] if history is not None else (forced.shape[1] if forced is not None else 0)
    if history is not None and forced is not None and forced.shape[1] != history.shape[1]:
        raise ValueError("argmax_advance: forced and history must have the same row length")
    check(lib.kx_argmax_advance(logits.data_ptr(), logits.stride(0), logits.shape[0], logits.shape[1], _ptr(forced),
                                tokens_out.data_ptr(), _ptr(history), hist_ld, _ptr(pos), step.data_ptr(),
                                counter.data_ptr(), _ptr(keys), _stream()), "kx_argmax_advance")


def _ptr_array(tensors):
    return (_abi.C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def decode_step_buffers(layers, device):
    """(plan bytes, barrier 288 x int64 (zeroed)) for kx_decode_plan_build."""
    return (torch.empty(int(lib.kx_decode_plan_bytes(layers)), dtype=torch.uint8, device=device),
            torch.zeros(288, dtype=torch.int64, device=device))


def decode_plan_build(plan, *, layers, out, embed_table, pos_table, tabs, k_cache, v_cache, tokens, x, xb, q, att, mid, logits,
                      keys, pos, step, err_flag, barrier, heads, ffn, t_max, eps, scale, forced=None, history=None, trace=None,
                      text_index_off=-1):
    """layers: list of dicts with (w, c, d) triples under "qkv", "o", "fc1", "fc2"; out: the (w, c, d) of the LM head."""
    g = _abi.DecodeStepArgs()
    B, D = x.shape
    g.batch, g.layers, g.d_model, g.ffn, g.heads, g.vocab, g.t_max = B, len(layers), D, ffn, heads, out[0].shape[0], t_max
    g.pos_rows, g.eps, g.scale = pos_table.shape[0], eps, scale
    g.text_index_off = int(text_index_off)
    keep = []
    for name in ("qkv", "o", "fc1", "fc2"):
        for j, letter in enumerate("wcd"):
            arr = _ptr_array([L[name][j] for L in layers])
            keep.append(arr)
            setattr(g, f"{letter}_{name}", _abi.C.cast(arr, _abi._pp))
    ka, va = _ptr_array(list(k_cache)), _ptr_array(list(v_cache))
    g.k_cache, g.v_cache = _abi.C.cast(ka, _abi._pp), _abi.C.cast(va, _abi._pp)
    g.w_out, g.c_out, g.d_out = out[0].data_ptr(), out[1].data_ptr(), _ptr(out[2])
    g.embed_table, g.pos_table = embed_table.data_ptr(), pos_table.data_ptr()
    g.xq_cos, g.xq_sin, g.xk_cos, g.xk_sin = (t.data_ptr() for t in tabs)
    g.tokens, g.x, g.xb, g.q, g.att, g.mid = (t.data_ptr() for t in (tokens, x, xb, q, att, mid))
    g.logits, g.ld_logits = logits.data_ptr(), logits.stride(0)
    g.argmax_keys, g.pos, g.step, g.err_flag = keys.data_ptr(), pos.data_ptr(), step.data_ptr(), _ptr(err_flag)
    g.forced, g.history = _ptr(forced), _ptr(history)
    g.history_ld = history.shape[1] if history is not None else (forced.shape[1] if forced is not None else 0)
    g.barrier = barrier.data_ptr()
    g.trace = _ptr(trace)
    check(lib.kx_decode_plan_build(g, plan.data_ptr(), _stream()), "kx_decode_plan_build")
    return plan


def decode_step(plan):
    """One new token per sequence: one persistent cooperative kernel (kx_decode_step)."""
    with _Timed("decode_step", 0.0, 0.0):
        check(lib.kx_decode_step(plan.data_ptr(), _stream()), "kx_decode_step")


_graph_launches = 0


def count_graph_replay(n: int) -> None:
    """A captured CUDA graph holding `n` of this library's kernels was replayed."""
    global _graph_launches
    _graph_launches += n


def launch_count() -> int:
    """Kernels of libkosmosx_sm100.so launched so far: direct launches (counted inside the library)
    plus kernels replayed from captured graphs."""
    return int(lib.kx_launch_count()) + _graph_launches
