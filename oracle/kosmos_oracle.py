"""CPU oracle for the Kosmos-X multimodal forward path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain eager PyTorch (fp32), the algorithm that
``kosmosx.model.Kosmos.forward(text_tokens, images)`` executes in the reference
(/root/reference/kosmosx/model.py:208-253).  It exists so that the CUDA path in
``kosmos-x_b200/`` can be checked against the reference semantics; it is never
imported by the product package.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s baseline legs (``cpu_baseline``, ``gpu_eager_baseline``, ``--impl reference``) may import it.

PARITY UNPINNED: the reference's hot-path arithmetic lives in
three third-party packages that are NOT vendored in /root/reference and are NOT
installed here (torchscale [unpinned, + the private ``passed_x`` patch of
README.md:179-193], flamingo_pytorch [unpinned], bitsandbytes 0.38.1), and the
reference's own tests hold no valid golden vectors for this path.
What *is* pinned:
  * the CLIP ViT-L/14 tower restated here is checked weight-for-weight against
    the installed ``transformers`` implementation (tests/test_oracle.py), which
    is the real dependency the reference calls at model.py:154-156,230;
  * the sub-LN decoder block and the whole decoder stack (xPos off) are checked against the installed
    ``transformers`` Kosmos-2 text block / ``Kosmos2TextForCausalLM``, an independent port of the same
    torchscale decoder (same parameter names);
  * xPos is checked for its defining relative-position (Toeplitz) property and, elementwise, against the
    xPos of the installed ``flash_attn`` (``RotaryEmbedding(scale_base=512, interleaved=True)``), an
    independent implementation of the same paper;
  * the perceiver attention and the whole resampler loop are checked against the installed
    ``transformers`` Idefics port of the same flamingo-pytorch modules (weights copied in; the port's ReLU
    swapped for flamingo's GELU, media position embedding zeroed);
  * ``clip_preprocess_u8`` equals ``transformers.image_transforms.rescale`` + ``normalize`` bit for bit.
Everything else follows the published algorithms of those packages as recorded
in SURVEY.md Appendix A; each class cites the reference call site it serves.

``emulate_bf16=True`` re-runs the same algorithm with every tensor-core operand
rounded to bf16 at exactly the points where the CUDA path rounds (fp32
accumulate, fp32 residual stream, fp32 LayerNorm/softmax/GELU), including the
LayerNorm fold of SURVEY.md A.7 (``_linear`` on an ``_LNOut``: statistics of the
bf16-rounded rows, gamma rounded into the weight).  It is used by the tests to
separate "kernel bug" from "bf16 rounding".
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
@dataclass
class OracleConfig:
    """Sizes of the reference model (defaults = /root/reference/kosmosx/model.py:149-206)."""

    # decoder (model.py:161-183)
    vocab: int = 32002
    dim: int = 2048
    layers: int = 24
    ffn: int = 8192
    heads: int = 32
    max_positions: int = 2048          # PositionalEmbedding(2048, 2048, 1): rows; T <= rows-2
    multiway: bool = True
    xpos_scale_base: int = 512
    eps: float = 1e-5
    # CLIP ViT-L/14 (model.py:154-156)
    vit_dim: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    patch: int = 14
    image: int = 224
    vit_act: str = "gelu"              # laion ViT-L/14 checkpoint; HF config default is quick_gelu
    # PerceiverResampler (model.py:196-203)
    p_depth: int = 2
    p_heads: int = 8
    p_dim_head: int = 64
    p_latents: int = 64
    p_media_embeds: int = 257
    p_ff_mult: int = 4
    # torchscale decoder.py, Decoder.forward_embedding [recall]: `x = embed = self.embed_scale * token_embedding` followed by
    # the IN-PLACE `x += positions` — `embed` (the `[1]` result the reference splices at model.py:238) is the same tensor
    # and carries the positions.  False = the out-of-place reading (`x = x + positions`).
    alias_embed_positions: bool = True
    # torchscale DecoderConfig(dropout=0.1, attention_dropout=0.1) at model.py:175-177 (training mode only)
    dropout: float = 0.1
    attention_dropout: float = 0.1

    @property
    def vit_tokens(self) -> int:
        return (self.image // self.patch) ** 2 + 1

    @staticmethod
    def tiny(**kw) -> "OracleConfig":
        """A small configuration with the same structure (head_dim stays 64)."""
        base = dict(vocab=1002, dim=128, layers=2, ffn=256, heads=2, max_positions=256,
                    vit_dim=128, vit_layers=2, vit_heads=2, vit_mlp=256, patch=14, image=56,
                    p_depth=2, p_heads=2, p_dim_head=64, p_latents=64, p_media_embeds=17)
        base.update(kw)
        return OracleConfig(**base)


class _Emu:
    """bf16 operand rounding switch shared by every oracle module of one model.  ``fold`` (with ``on``): LayerNorms that
    the CUDA path folds into their consumer GEMM (every decoder LayerNorm, the ViT's layer_norm1/2) are
    evaluated the way the kernels evaluate them, see ``_linear``."""

    def __init__(self, on: bool = False, fold: bool = True):
        self.on = on
        self.fold = fold
        # training-mode dropout with INJECTED masks (multipliers: 0 or 1 / keep), so that a parity test can hand the
        # oracle exactly the masks the CUDA path drew: {"x0": (B,T,D), ("attn_out", layer): (B,T,D),
        # ("ffn_out", layer): (B,T,D), ("attn", layer): (B,H,T,T)}; None = eval mode (identity)
        self.drop = None

    def d(self, x, key):
        if self.drop is None or key not in self.drop:
            return x
        return x * self.drop[key].to(x.dtype).reshape(x.shape)

    def r(self, x: torch.Tensor) -> torch.Tensor:
        return x.to(torch.bfloat16).to(torch.float32) if self.on else x


class _LNOut:
    """A LayerNorm whose evaluation is deferred to the Linear that consumes it (bf16 emulation of the folded path)."""

    def __init__(self, x, ln):
        self.x, self.ln = x, ln


def _ln(emu: _Emu, ln: nn.LayerNorm, x):
    """ln(x) — or, when emulating the folded kernels, the raw rows + the LayerNorm for ``_linear`` to fold."""
    if emu.on and emu.fold:
        return _LNOut(x, ln)
    return ln(x)


def _linear(emu: _Emu, x, lin: nn.Linear):
    if isinstance(x, _LNOut):
        # what kx_gemm_bf16 computes with kx_gemm_args.ln_part (gemm.cu: ln_row_stats, epilogue_math; model.py: _fold_ln):
        #   xb = bf16(x);  mean, E[x^2] - mean^2 of the ROUNDED row;  W' = bf16(W * gamma);  c = W'.1;  d = W.beta + b
        #   y = rstd * (xb . W'^T - mean * c) + d
        ln = x.ln
        xb = emu.r(x.x)
        n = xb.shape[-1]
        mean = xb.sum(-1, keepdim=True) / n
        var = ((xb * xb).sum(-1, keepdim=True) / n - mean * mean).clamp_min(0.0)
        rstd = torch.rsqrt(var + ln.eps)
        wp = emu.r(lin.weight * ln.weight[None, :])
        c = wp.double().sum(1).float()
        d = lin.weight.double() @ ln.bias.double()
        if lin.bias is not None:
            d = d + lin.bias.double()
        return rstd * (F.linear(xb, wp) - mean * c) + d.float()
    return F.linear(emu.r(x), emu.r(lin.weight), lin.bias)


# --------------------------------------------------------------------------- CLIP ViT
class _ClipAttention(nn.Module):
    """[HF] transformers/models/clip/modeling_clip.py:282-336 (eager attention)."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        d = cfg.vit_dim
        self.h = cfg.vit_heads
        self.emu = emu
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)

    def forward(self, x):
        B, T, D = (x.x if isinstance(x, _LNOut) else x).shape
        hd = D // self.h
        e = self.emu
        q = e.r(_linear(e, x, self.q_proj) * hd ** -0.5).view(B, T, self.h, hd).transpose(1, 2)
        k = e.r(_linear(e, x, self.k_proj)).view(B, T, self.h, hd).transpose(1, 2)
        v = e.r(_linear(e, x, self.v_proj)).view(B, T, self.h, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        m = s.amax(-1, keepdim=True)
        p = torch.exp(s - m)
        o = (e.r(p) @ v) / p.sum(-1, keepdim=True)
        o = e.r(o).transpose(1, 2).reshape(B, T, D)
        return _linear(e, o, self.out_proj)


class _ClipMLP(nn.Module):
    """[HF] modeling_clip.py:339-351."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        self.emu = emu
        self.act = cfg.vit_act
        self.fc1 = nn.Linear(cfg.vit_dim, cfg.vit_mlp)
        self.fc2 = nn.Linear(cfg.vit_mlp, cfg.vit_dim)

    def forward(self, x):
        h = _linear(self.emu, x, self.fc1)
        h = F.gelu(h) if self.act == "gelu" else h * torch.sigmoid(1.702 * h)
        return _linear(self.emu, h, self.fc2)


class _ClipLayer(nn.Module):
    """[HF] modeling_clip.py:363-385 (pre-LN block)."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        self.self_attn = _ClipAttention(cfg, emu)
        self.layer_norm1 = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)
        self.mlp = _ClipMLP(cfg, emu)
        self.layer_norm2 = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)

    def forward(self, x):
        e = self.self_attn.emu
        x = x + self.self_attn(_ln(e, self.layer_norm1, x))
        return x + self.mlp(_ln(e, self.layer_norm2, x))


class _ClipEmbeddings(nn.Module):
    """[HF] modeling_clip.py:138-159, 202-218."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        self.cfg, self.emu = cfg, emu
        self.class_embedding = nn.Parameter(torch.randn(cfg.vit_dim))
        self.patch_embedding = nn.Conv2d(3, cfg.vit_dim, cfg.patch, cfg.patch, bias=False)
        self.position_embedding = nn.Embedding(cfg.vit_tokens, cfg.vit_dim)

    def forward(self, pixels):
        B, _, H, W = pixels.shape
        if H != self.cfg.image or W != self.cfg.image:
            raise ValueError(f"Input image size ({H}*{W}) doesn't match model "
                             f"({self.cfg.image}*{self.cfg.image}).")
        e = self.emu
        w = self.patch_embedding.weight
        x = F.conv2d(e.r(pixels.to(w.dtype)), e.r(w), stride=self.cfg.patch)
        x = x.flatten(2).transpose(1, 2)
        x = torch.cat([self.class_embedding.expand(B, 1, -1), x], dim=1)
        return x + self.position_embedding.weight[None]


class _ClipEncoder(nn.Module):
    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        self.layers = nn.ModuleList(_ClipLayer(cfg, emu) for _ in range(cfg.vit_layers))


class ClipVisionTower(nn.Module):
    """CLIPVisionTransformer, [HF] modeling_clip.py:647-697.  ``forward`` returns the
    un-normalised ``last_hidden_state`` that model.py:230 consumes (post_layernorm is
    applied to the pooled CLS token only and is unused on this path)."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        self.embeddings = _ClipEmbeddings(cfg, emu)
        self.pre_layrnorm = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)   # (sic) HF attribute name
        self.encoder = _ClipEncoder(cfg, emu)
        self.post_layernorm = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)

    def forward(self, pixel_values):
        x = self.pre_layrnorm(self.embeddings(pixel_values))
        for layer in self.encoder.layers:
            x = layer(x)
        return x


# --------------------------------------------------------------------------- perceiver
class _PerceiverAttention(nn.Module):
    """flamingo_pytorch PerceiverAttention; call site model.py:196-203,231."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        d, inner = cfg.vit_dim, cfg.p_heads * cfg.p_dim_head
        self.h, self.dh, self.emu = cfg.p_heads, cfg.p_dim_head, emu
        self.norm_media = nn.LayerNorm(d)
        self.norm_latents = nn.LayerNorm(d)
        self.to_q = nn.Linear(d, inner, bias=False)
        self.to_kv = nn.Linear(d, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, d, bias=False)

    def forward(self, x, latents):
        # x: (B, m, n, D)   latents: (B, m, L, D)
        e = self.emu
        x = self.norm_media(x)
        latents = self.norm_latents(latents)
        B, m, L, _ = latents.shape
        q = _linear(e, latents, self.to_q)
        kv = _linear(e, torch.cat([x, latents], dim=-2), self.to_kv)
        k, v = kv.chunk(2, dim=-1)

        def split(t):                                      # b m n (h d) -> b h m n d
            return t.view(B, m, t.shape[2], self.h, self.dh).permute(0, 3, 1, 2, 4)

        q, k, v = e.r(split(q)), e.r(split(k)), e.r(split(v))
        q = q * self.dh ** -0.5
        sim = q @ k.transpose(-1, -2)
        sim = sim - sim.amax(dim=-1, keepdim=True)
        attn = sim.softmax(dim=-1)
        out = e.r(attn @ v)                                # b h m L d
        out = out.permute(0, 2, 3, 1, 4).reshape(B, m, L, self.h * self.dh)
        return _linear(e, out, self.to_out)


class _EmuLinear(nn.Linear):
    """nn.Linear whose operands go through the bf16 emulation switch (keeps the
    ``Sequential`` index names ``1``/``3`` of the flamingo FeedForward)."""

    emu: _Emu = None

    def forward(self, x):
        return _linear(self.emu, x, self)


def _perceiver_ff(cfg: OracleConfig, emu: _Emu):
    d, inner = cfg.vit_dim, cfg.vit_dim * cfg.p_ff_mult
    l1, l3 = _EmuLinear(d, inner, bias=False), _EmuLinear(inner, d, bias=False)
    l1.emu = l3.emu = emu
    return nn.Sequential(nn.LayerNorm(d), l1, nn.GELU(), l3)


class PerceiverResampler(nn.Module):
    """flamingo_pytorch.PerceiverResampler(dim=1024, depth=2, dim_head=64, heads=8,
    num_latents=64, num_media_embeds=257) — SURVEY.md A.2; model.py:196-203."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        d = cfg.vit_dim
        self.latents = nn.Parameter(torch.randn(cfg.p_latents, d))
        self.media_pos_emb = nn.Parameter(torch.randn(cfg.p_media_embeds, 1, d))
        self.layers = nn.ModuleList(
            nn.ModuleList([_PerceiverAttention(cfg, emu), _perceiver_ff(cfg, emu)])
            for _ in range(cfg.p_depth))
        self.norm = nn.LayerNorm(d)

    def forward(self, x):
        if x.ndim == 3:
            x = x[:, None]                                  # (B, 1, n, D)
        m = x.shape[1]
        x = x + self.media_pos_emb[:m]                      # indexed by MEDIA index (A.2)
        latents = self.latents.expand(x.shape[0], m, -1, -1)
        for attn, ff in self.layers:
            latents = attn(x, latents) + latents
            latents = ff(latents) + latents
        return self.norm(latents)


# --------------------------------------------------------------------------- xPos
def _rotate_every_two(x):
    x1, x2 = x[..., ::2], x[..., 1::2]
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def _dup_interleave(m):
    return m.repeat_interleave(2, dim=-1)


class XPOS(nn.Module):
    """torchscale.component.xpos_relative_position.XPOS; enabled by
    ``xpos_rel_pos=True`` at model.py:180."""

    def __init__(self, head_dim: int, scale_base: int = 512):
        super().__init__()
        self.head_dim, self.scale_base = head_dim, scale_base
        self.register_buffer("scale", (torch.arange(0, head_dim, 2) + 0.4 * head_dim) / (1.4 * head_dim))

    def tables(self, length: int, offset: int = 0):
        """(scale[T,hd/2], sin[T,hd/2], cos[T,hd/2]) before the up/down choice."""
        min_pos = -(length + offset) // 2
        max_pos = length + offset + min_pos
        scale = self.scale ** (torch.arange(min_pos, max_pos, 1).to(self.scale) / self.scale_base)[:, None]
        dim = scale.shape[1]
        inv_freq = 1.0 / (10000 ** (torch.arange(0, dim) / dim))
        sinusoid = torch.einsum("i , j -> i j", torch.arange(0, scale.shape[0], dtype=torch.float), inv_freq).to(scale)
        return scale, torch.sin(sinusoid), torch.cos(sinusoid)

    def forward(self, x, offset=0, downscale=False):
        length = x.shape[1]
        scale, sin, cos = self.tables(length, offset)
        if scale.shape[0] > length:
            scale, sin, cos = scale[-length:], sin[-length:], cos[-length:]
        if downscale:
            scale = 1 / scale
        sin, cos = _dup_interleave(sin * scale), _dup_interleave(cos * scale)
        return x * cos + _rotate_every_two(x) * sin


# --------------------------------------------------------------------------- decoder
class MultiwayNetwork(nn.Module):
    """torchscale.component.multiway_network.MultiwayNetwork.  Nothing in
    the reference sets ``split_position`` so only ``.A`` ever runs (model.py:181)."""

    def __init__(self, module: nn.Module, make_b):
        super().__init__()
        self.A = module
        self.B = make_b()
        self.split_position = -1

    def forward(self, x, **kw):
        if self.split_position == -1:
            return self.A(x, **kw)
        if self.split_position == 0:
            return self.B(x, **kw)
        x1, x2 = torch.split(x, [self.split_position, x.size(1) - self.split_position], dim=1)
        return torch.cat([self.A(x1, **kw), self.B(x2, **kw)], dim=1)


def _wrap(cfg: OracleConfig, make):
    return MultiwayNetwork(make(), make) if cfg.multiway else make()


def _inner(mod):
    """The live branch of a (possibly) multiway-wrapped module."""
    return mod.A if isinstance(mod, MultiwayNetwork) else mod


class FeedForwardNetwork(nn.Module):
    """torchscale FeedForwardNetwork with sub-LN."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        self.emu = emu
        self.fc1 = nn.Linear(cfg.dim, cfg.ffn)
        self.fc2 = nn.Linear(cfg.ffn, cfg.dim)
        self.ffn_layernorm = nn.LayerNorm(cfg.ffn, eps=cfg.eps)

    def forward(self, x):
        x = _linear(self.emu, x, self.fc1)
        x = F.gelu(x.float()).type_as(x)
        x = _ln(self.emu, self.ffn_layernorm, self.emu.r(x))
        return _linear(self.emu, x, self.fc2)


class MultiheadAttention(nn.Module):
    """torchscale MultiheadAttention, self-attention with sub-LN and xPos."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        d = cfg.dim
        self.h, self.hd, self.emu = cfg.heads, cfg.dim // cfg.heads, emu
        self.scaling = self.hd ** -0.5
        self.k_proj = _wrap(cfg, lambda: nn.Linear(d, d))
        self.v_proj = _wrap(cfg, lambda: nn.Linear(d, d))
        self.q_proj = _wrap(cfg, lambda: nn.Linear(d, d))
        self.out_proj = _wrap(cfg, lambda: nn.Linear(d, d))
        self.inner_attn_ln = _wrap(cfg, lambda: nn.LayerNorm(d, eps=cfg.eps))
        self.xpos = XPOS(self.hd, cfg.xpos_scale_base)
        self.use_xpos = True            # tests switch it off to cross-check against HF Kosmos-2's block
        self.index = 0                  # layer index (set by Decoder): names this layer's dropout masks

    def forward(self, x, attn_mask, incremental_state=None, is_first_step=False):
        """``incremental_state`` (a per-layer dict) follows torchscale's MultiheadAttention [recall]: the UN-rotated
        keys / values of every step are appended to ``prev_key`` / ``prev_value`` (B, heads, src_len, head_dim) and
        xPos is re-applied to the whole key sequence each step; after the first step the single query is rotated
        with ``offset = src_len - 1``."""
        B, T, D = (x.x if isinstance(x, _LNOut) else x).shape
        e = self.emu
        q = _linear(e, x, _inner(self.q_proj))
        k = _linear(e, x, _inner(self.k_proj))
        v = _linear(e, x, _inner(self.v_proj))
        q = q * self.scaling

        def heads(t):
            return t.view(B, T, self.h, self.hd).transpose(1, 2).reshape(B * self.h, T, self.hd)

        q, k, v = heads(q), heads(k), e.r(heads(v))
        offset = 0
        if incremental_state is not None:
            if "prev_key" in incremental_state:
                k = torch.cat([incremental_state["prev_key"].view(B * self.h, -1, self.hd), k], dim=1)
                v = torch.cat([incremental_state["prev_value"].view(B * self.h, -1, self.hd), v], dim=1)
            incremental_state["prev_key"] = k.view(B, self.h, -1, self.hd)
            incremental_state["prev_value"] = v.view(B, self.h, -1, self.hd)
            if not is_first_step:
                offset = k.size(1) - 1
        if self.use_xpos:
            k = self.xpos(k, offset=0, downscale=True)
            q = self.xpos(q, offset=offset, downscale=False)
        k, q = e.r(k), e.r(q)
        w = torch.bmm(q, k.transpose(1, 2))
        w = torch.nan_to_num(w)
        if attn_mask is not None:
            w = w + attn_mask[None]
        m = w.amax(-1, keepdim=True)
        p = torch.exp(w - m)                                 # == softmax(w, dtype=fp32) numerator
        pd = e.d(p, ("attn", self.index))                    # attention dropout acts on the normalised probabilities
        a = torch.bmm(e.r(pd), v) / p.sum(-1, keepdim=True)
        a = e.r(a).view(B, self.h, T, self.hd).transpose(1, 2).reshape(B, T, D)
        a = _ln(e, _inner(self.inner_attn_ln), a)
        return _linear(e, a, _inner(self.out_proj))


class DecoderLayer(nn.Module):
    """torchscale DecoderLayer, pre-LN (forced by subln), alpha = 1."""

    def __init__(self, cfg: OracleConfig, emu: _Emu):
        super().__init__()
        self.self_attn = MultiheadAttention(cfg, emu)
        self.self_attn_layer_norm = _wrap(cfg, lambda: nn.LayerNorm(cfg.dim, eps=cfg.eps))
        self.ffn = _wrap(cfg, lambda: FeedForwardNetwork(cfg, emu))
        self.final_layer_norm = _wrap(cfg, lambda: nn.LayerNorm(cfg.dim, eps=cfg.eps))

    def forward(self, x, mask, incremental_state=None, is_first_step=False):
        r = x
        e = self.self_attn.emu
        x = self.self_attn(_ln(e, _inner(self.self_attn_layer_norm), x), mask, incremental_state, is_first_step)
        x = r + e.d(x, ("attn_out", self.self_attn.index))   # x = dropout(self_attn(...)); x = residual + x
        r = x
        x = _inner(self.ffn)(_ln(e, _inner(self.final_layer_norm), x))
        return r + e.d(x, ("ffn_out", self.self_attn.index)) # FeedForwardNetwork ends with dropout(fc2(...))


class PositionalEmbedding(nn.Embedding):
    """torchscale.component.embedding.PositionalEmbedding: positions start
    at 2 ("consistent with fairseq"); only ``x.size(1)`` of the argument is used."""

    def forward(self, x, positions=None, **kwargs):
        if positions is None:
            positions = torch.arange(2, x.size(1) + 2, device=x.device).long().unsqueeze(0)
        return F.embedding(positions, self.weight, self.padding_idx)


class Decoder(nn.Module):
    """torchscale.architecture.decoder.Decoder + the ``passed_x`` patch
    (/root/reference/README.md:179-193); call sites model.py:186-191,238,242,250."""

    def __init__(self, cfg: OracleConfig, emu: _Emu, embed_tokens, embed_positions, output_projection):
        super().__init__()
        self.cfg, self.emu = cfg, emu
        self.embed_scale = 1.0                               # no_scale_embedding=True
        self.embed_tokens = embed_tokens
        self.embed_positions = embed_positions
        self.output_projection = output_projection
        self.layers = nn.ModuleList(DecoderLayer(cfg, emu) for _ in range(cfg.layers))
        for i, layer in enumerate(self.layers):
            layer.self_attn.index = i
        self.layer_norm = nn.LayerNorm(cfg.dim, eps=cfg.eps)
        # sub-LN init (decoder-only): scale sqrt(log(2*layers)) on fc1/fc2/out_proj/v_proj
        init_scale = math.sqrt(math.log(cfg.layers * 2))
        for name, p in self.named_parameters():
            if "fc1" in name or "fc2" in name or "out_proj" in name or "v_proj" in name:
                p.data.mul_(init_scale)

    @staticmethod
    def is_first_step(incremental_state):
        if incremental_state is None:
            return False
        return incremental_state.get("is_first_step", False)

    def forward_embedding(self, tokens, token_embedding=None, incremental_state=None):
        positions = self.embed_positions(tokens, incremental_state=incremental_state)
        if incremental_state is not None and not self.is_first_step(incremental_state):
            # later decoding steps: the caller passes the whole prefix (fairseq style), only its last
            # token is embedded and it takes the last position, len + 1  [recall: torchscale decoder.py]
            tokens = tokens[:, -1:]
            positions = positions[:, -1:]
            if token_embedding is not None:
                token_embedding = token_embedding[:, -1:]
        if token_embedding is None:
            token_embedding = self.embed_tokens(tokens)
        x = embed = self.embed_scale * token_embedding
        if self.cfg.alias_embed_positions:
            x += positions                                   # in place: `embed` IS x (torchscale as written)
        else:
            x = x + positions
        return x, embed                                      # dropout p=0.1 is identity in eval

    def forward(self, prev_output_tokens, incremental_state=None, token_embeddings=None, **kwargs):
        """``incremental_state``: torchscale's generation protocol [recall] — a dict the caller creates as
        ``{"is_first_step": True}`` for the prompt pass (all positions run, causal mask, K/V cached per layer under
        integer keys) and flips to ``False`` for the following one-token steps (no mask, K/V appended)."""
        if kwargs.get("passed_x", None) is None:
            x, _ = self.forward_embedding(prev_output_tokens, token_embeddings, incremental_state)
        else:
            x = kwargs["passed_x"]
        first = self.is_first_step(incremental_state)
        T = x.size(1)
        inner_states = [x]
        for idx, layer in enumerate(self.layers):
            if incremental_state is None or first:
                mask = torch.triu(torch.zeros([T, T]).float().fill_(float("-inf")).type_as(x), 1)
            else:
                mask = None
            st = None
            if incremental_state is not None:
                st = incremental_state.setdefault(idx, {})
            x = layer(x, mask, st, first)
            inner_states.append(x)
        x = _linear(self.emu, _ln(self.emu, self.layer_norm, x), self.output_projection)
        return x, {"inner_states": inner_states, "l_aux": [None] * len(self.layers), "attn": None}


# --------------------------------------------------------------------------- top level
def _embed_init(emb: nn.Embedding):
    """bitsandbytes.nn.modules.Embedding.reset_parameters: xavier-uniform, padding row zero."""
    nn.init.xavier_uniform_(emb.weight)
    with torch.no_grad():
        emb.weight[emb.padding_idx].fill_(0)


class KosmosOracle(nn.Module):
    """``kosmosx.model.Kosmos`` (/root/reference/kosmosx/model.py:132-253), random-init CLIP."""

    def __init__(self, cfg: OracleConfig | None = None, emulate_bf16: bool = False):
        super().__init__()
        cfg = cfg or OracleConfig()
        self.cfg = cfg
        self.emu = _Emu(emulate_bf16)
        self.clip_model = ClipVisionTower(cfg, self.emu)                       # model.py:154-156
        self.embed = nn.Embedding(cfg.vocab, cfg.dim, padding_idx=1)           # model.py:161-163
        _embed_init(self.embed)
        self.embed_positions = PositionalEmbedding(cfg.max_positions, cfg.dim, 1)   # model.py:164
        self.output_projection = nn.Linear(cfg.dim, cfg.vocab, bias=False)     # model.py:166-167
        nn.init.normal_(self.output_projection.weight, mean=0, std=cfg.dim ** -0.5)
        self.decoder = Decoder(cfg, self.emu, self.embed, self.embed_positions, self.output_projection)
        self.perceive = PerceiverResampler(cfg, self.emu)                      # model.py:196-203
        self.image_proj = nn.Linear(cfg.vit_dim, cfg.dim, bias=False)          # model.py:205-206
        nn.init.normal_(self.image_proj.weight, mean=0, std=cfg.dim ** -0.5)

    def set_emulation(self, on: bool, fold: bool = True):
        self.emu.on = on
        self.emu.fold = fold

    def set_dropout_masks(self, masks: dict | None):
        """Training-mode dropout with the given multiplier tensors (see _Emu.drop); None switches it off."""
        self.emu.drop = masks

    def _image_rows(self, images, keep=None):
        """ViT -> perceiver -> image_proj: (B,3,H,W) -> (B,1,64,dim); (B,m,3,H,W) -> (B,m,64,dim)."""
        feats = self._resample(images)
        if keep is not None:
            keep["perceive"] = feats.squeeze(1) if images.ndim == 4 else feats
        return F.linear(self.emu.r(feats), self.emu.r(self.image_proj.weight)) # model.py:232

    def _resample(self, images):
        if images.ndim == 5:                                                   # config 5: m images per sequence
            B, m = images.shape[:2]
            feats = self.clip_model(pixel_values=images.flatten(0, 1))         # model.py:230 per image
            feats = self.perceive(feats.view(B, m, *feats.shape[1:]))          # (B, m, 64, Dv): media index = image index (A.2)
        else:
            feats = self.perceive(self.clip_model(pixel_values=images))        # model.py:230-231 (B, 1, 64, Dv)
        return feats

    @staticmethod
    def splice(embed, rows, image_positions):
        """torch.cat of model.py:239-241, generalised: image i's rows go in front of text token
        image_positions[i] (the reference is the single-image case image_positions = [2])."""
        m = rows.shape[1]
        pos = [2] if image_positions is None else list(image_positions)
        if len(pos) != m or sorted(pos) != pos or pos[0] < 0 or pos[-1] > embed.shape[1]:
            raise ValueError(f"image_positions {pos} does not describe {m} ascending text positions")
        parts, prev = [], 0
        for i, p in enumerate(pos):
            parts += [embed[:, prev:p], rows[:, i]]
            prev = p
        parts.append(embed[:, prev:])
        return torch.cat(parts, dim=1)

    def embed_inputs(self, text_tokens, images, image_positions=None):
        """model.py:230-244: the decoder input x0 (B, T_text + 64*m, dim), positions added."""
        rows = self._image_rows(images)
        model_input = self.decoder.forward_embedding(text_tokens)[1]           # model.py:238
        model_input = self.splice(model_input, rows, image_positions)          # model.py:239-241
        x0 = self.decoder.forward_embedding(model_input, token_embedding=model_input)[0]   # model.py:242-244
        return self.emu.d(x0, "x0")        # forward_embedding ends with dropout(x) (training mode; masks injected, see _Emu)

    def forward(self, text_tokens, images, image_positions=None, **kwargs):
        if not isinstance(text_tokens, torch.Tensor) or not isinstance(images, torch.Tensor):
            raise TypeError("text_tokens and images must be instances of torch.Tensor")
        model_input = self.embed_inputs(text_tokens, images, image_positions)
        return self.decoder(model_input, passed_x=model_input)[0]              # model.py:250

    @torch.no_grad()
    def generate(self, text_tokens, images, max_new_tokens, image_positions=None, forced=None):
        """Greedy continuation through torchscale's incremental protocol: the prompt pass runs
        ``decoder(x, incremental_state={"is_first_step": True}, passed_x=x)`` on the spliced input of model.py:230-244,
        every later step ``decoder(prefix, incremental_state=st)`` with only the prefix's last token embedded at
        position ``len(prefix) + 1``.  ``forced`` (B, n) replaces the argmax choice (teacher forcing for parity
        tests).  Returns (tokens (B, n), logits (B, n, vocab)); logits[:, i] is the distribution token i was drawn from."""
        x = self.embed_inputs(text_tokens, images, image_positions)
        st = {"is_first_step": True}
        logits = self.decoder(x, incremental_state=st, passed_x=x)[0][:, -1]
        st["is_first_step"] = False
        prefix = torch.zeros(x.size(0), x.size(1), dtype=torch.long)      # only its length and last token are read
        toks, outs = [], []
        t_text = text_tokens.shape[1]
        for i in range(max_new_tokens):
            nxt = logits.argmax(-1) if forced is None else forced[:, i]
            toks.append(nxt)
            outs.append(logits)
            if i + 1 < max_new_tokens:
                prefix = torch.cat([prefix, nxt[:, None]], dim=1)
                te = None
                if self.cfg.alias_embed_positions:
                    # the reference has no generate: a user would call forward() on the grown text, where text token
                    # t_text + i gets its text-index position as well (aliased `embed`); forward_embedding adds the other
                    te = torch.zeros(x.size(0), prefix.size(1), self.cfg.dim)
                    te[:, -1] = self.embed(nxt) + self.embed_positions.weight[t_text + i + 2]
                logits = self.decoder(prefix, incremental_state=st, token_embeddings=te)[0][:, -1]
        return torch.stack(toks, 1), torch.stack(outs, 1)

    @staticmethod
    def loss_targets(text_tokens, n_latents, image_positions=None, n_images=1, rule="reference", pad_token_id=None):
        """Next-token targets per row of the spliced sequence, -100 = no loss.

        rule = "reference": the reference's intended loop, experimental/model/allModalities/notes.txt:566-574 —
        ``outputs = cat([outputs[:, :1], outputs[:, 67:]])`` then ``loss(outputs[:, :-1], labels[:, 1:])`` with
        ``labels`` = the text WITHOUT the ``<image>`` ``</image>`` markers (model.py:70-77): rows 0 and 67.. are kept,
        row 0 predicts the first real text token (text token 3), the marker rows and the 64 feature rows carry no
        loss.  Generalised to any image position p: text tokens p-1 and p are that image's markers.  The class
        dimension is the vocabulary (the pasted loop hands CrossEntropyLoss a (B, T, C) tensor, classes on dim 1 — a
        defect).
        rule = "next_token": the row of text token i predicts text token i+1; feature rows, the last token and the
        token directly in front of an image carry no loss.
        pad_token_id: targets equal to it are ignored (the reference's loop masks nothing: None)."""
        B, t_text = text_tokens.shape
        pos = [2] * n_images if image_positions is None else list(image_positions)
        T = t_text + n_latents * len(pos)
        starts = [p + i * n_latents for i, p in enumerate(pos)]
        is_img = torch.zeros(T, dtype=torch.bool)
        for r in starts:
            is_img[r:r + n_latents] = True
        text_rows = (~is_img).nonzero().flatten().tolist()
        tgt = torch.full((B, T), -100, dtype=torch.long)
        if rule == "next_token":
            for ti, t in enumerate(text_rows):
                if ti + 1 < t_text and not (t + 1 < T and bool(is_img[t + 1])):
                    tgt[:, t] = text_tokens[:, ti + 1]
        elif rule == "reference":
            markers = {q for p in pos for q in (p - 1, p)}
            real = [ti for ti in range(t_text) if ti not in markers]       # `labels` = only_text_tokens (model.py:77)
            for a, b in zip(real[:-1], real[1:]):
                tgt[:, text_rows[a]] = text_tokens[:, b]
        else:
            raise ValueError("rule must be 'reference' or 'next_token'")
        if pad_token_id is not None:
            tgt[tgt == pad_token_id] = -100
        return tgt

    def loss(self, text_tokens, images, image_positions=None, rule="reference", pad_token_id=None):
        """Mean cross-entropy over the text rows (train.py:647 `loss = model(...)`; see loss_targets)."""
        logits = self.forward(text_tokens, images, image_positions=image_positions)
        m = images.shape[1] if images.ndim == 5 else 1
        tgt = self.loss_targets(text_tokens, self.cfg.p_latents, image_positions, m, rule, pad_token_id)
        return F.cross_entropy(logits.reshape(-1, logits.shape[-1]), tgt.reshape(-1), ignore_index=-100)

    @torch.no_grad()
    def stages(self, text_tokens, images, image_positions=None):
        """Intermediate tensors, for per-stage parity tests."""
        out = {}
        flat = images.flatten(0, 1) if images.ndim == 5 else images
        out["vit"] = self.clip_model(pixel_values=flat)
        rows = self._image_rows(images, keep=out)
        out["image_proj"] = rows.squeeze(1) if images.ndim == 4 else rows
        emb = self.decoder.forward_embedding(text_tokens)[1]
        x = self.splice(emb, rows, image_positions)
        out["x0"] = self.decoder.forward_embedding(x, token_embedding=x)[0]
        logits, extra = self.decoder(out["x0"], passed_x=out["x0"])
        out["inner_states"] = extra["inner_states"]
        out["logits"] = logits
        return out


class KosmosLanguageOracle(nn.Module):
    """``kosmosx.model.KosmosLanguage`` (/root/reference/kosmosx/model.py:256-320)."""

    def __init__(self, cfg: OracleConfig | None = None, emulate_bf16: bool = False):
        super().__init__()
        cfg = cfg or OracleConfig(vocab=64007, max_positions=2048)
        self.cfg = cfg
        self.emu = _Emu(emulate_bf16)
        self.embed = nn.Embedding(cfg.vocab, cfg.dim, padding_idx=1)
        _embed_init(self.embed)
        self.embed_positions = PositionalEmbedding(cfg.max_positions, cfg.dim, 1)
        self.output_projection = nn.Linear(cfg.dim, cfg.vocab, bias=False)
        self.decoder = Decoder(cfg, self.emu, self.embed, self.embed_positions, self.output_projection)

    def forward(self, x, **kwargs):
        model_input = self.decoder.forward_embedding(x, **kwargs)[0]
        return self.decoder(model_input, passed_x=model_input)[0]


# --------------------------------------------------------------------------- helpers
# --------------------------------------------------------------------------- host preprocessing
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)     # HF image_utils.py; the laion ViT-L/14 processor's values
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_preprocess_u8(pixels, channels_last: bool = False, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD) -> torch.Tensor:
    """``KosmosTokenizer.tokenize_images`` (reference model.py:82-97 -> ``CLIPProcessor``, transformers 4.35 slow
    image processor) for uint8 images that already have the model's size, so that resize / centre crop are
    identities: ``rescale`` = uint8 -> float64, * (1/255), -> float32; ``normalize`` = (x - mean) / std in float32
    (HF image_transforms.py ``rescale`` / ``normalize``; tests/test_oracle.py pins this function bit for bit
    against those two functions of the installed transformers, and to 1e-6 against ``CLIPImageProcessor``).
    pixels: uint8 (N,3,H,W), or (N,H,W,3) with ``channels_last``.  -> fp32 pixel_values (N,3,H,W)."""
    import numpy as np
    a = pixels.cpu().numpy() if isinstance(pixels, torch.Tensor) else np.asarray(pixels)
    assert a.dtype == np.uint8 and a.ndim == 4
    if channels_last:
        a = a.transpose(0, 3, 1, 2)
    x = (a.astype(np.float64) * (1 / 255)).astype(np.float32)
    m = np.array(mean, dtype=np.float32).reshape(1, 3, 1, 1)
    s = np.array(std, dtype=np.float32).reshape(1, 3, 1, 1)
    return torch.from_numpy(np.ascontiguousarray((x - m) / s))


def _pil_bicubic_axis(img, out_size: int, axis: int):
    """One pass of PIL's ImagingResample (8 bits per channel, BICUBIC) along `axis` of a uint8 (H, W, C) array: taps from
    ``precompute_coeffs`` (support 2 x max(scale, 1), weights normalised to 1 in float64), fixed-point
    ``normalize_coeffs_8bpc`` (2^22, half away from zero), accumulator seeded with 2^21, ``clip8(acc >> 22)``."""
    import numpy as np
    bits = 32 - 8 - 2
    in_size = img.shape[axis]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)

    def bicubic(x, a=-0.5):
        x = abs(x)
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0

    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [bicubic((x + xmin - center + 0.5) / filterscale) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        k = np.array([int(-0.5 + v * (1 << bits)) if v < 0 else int(0.5 + v * (1 << bits)) for v in w], dtype=np.int64)
        acc = np.tensordot(k, src[xmin:xmin + xmax], axes=(0, 0)) + (1 << (bits - 1))
        out[xx] = np.clip(acc >> bits, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def clip_resize_center_crop_u8(image, size: int = 224, crop: int | None = None):
    """``CLIPImageProcessor`` resize + centre crop of transformers 4.35 (the processor the reference builds at
    model.py:36-38 and calls at model.py:81-97) for ONE uint8 (H, W, 3) image: ``get_resize_output_image_size`` with
    ``shortest_edge`` (long side = int(size * long / short)), ``PIL.Image.resize(BICUBIC)`` (horizontal pass, uint8,
    vertical pass), ``center_crop`` (top = (h - crop) // 2).  Returns uint8 (crop, crop, 3).  tests/test_preprocess.py
    pins it bit for bit against transformers' own ``resize`` / ``center_crop`` (which call PIL)."""
    import numpy as np
    crop = size if crop is None else crop
    a = image.cpu().numpy() if isinstance(image, torch.Tensor) else np.asarray(image)
    assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3
    h, w = a.shape[:2]
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_h, new_w = (new_long, new_short) if w <= h else (new_short, new_long)
    if w != new_w:
        a = _pil_bicubic_axis(a, new_w, 1)
    if h != new_h:
        a = _pil_bicubic_axis(a, new_h, 0)
    top, left = (new_h - crop) // 2, (new_w - crop) // 2
    return torch.from_numpy(np.ascontiguousarray(a[top:top + crop, left:left + crop]))


def tokenize_texts(input_ids: torch.Tensor, im_idx: int, im_end_idx: int):
    """``KosmosTokenizer.tokenize_texts`` after the HF tokenizer call (reference model.py:68-76):
    ``<s> <image> </image> text </s>`` -> (tokens with the two image tokens after position 0, the text tokens)."""
    image_tokens = torch.tensor([[im_idx, im_end_idx]] * input_ids.shape[0])
    return torch.cat([input_ids[:, 0:1], image_tokens, input_ids[:, 1:]], dim=1), input_ids


def tokenize_attention_mask(text_tokens: torch.Tensor, pad_token_id: int, n_image_features: int = 64):
    """The mask of ``KosmosTokenizer.tokenize`` (reference model.py:113-120): 64 ones in front of (tokens != pad)."""
    attention_mask = text_tokens != pad_token_id
    return torch.cat([torch.ones((text_tokens.shape[0], n_image_features)), attention_mask], dim=1)


def make_inputs(cfg: OracleConfig, batch: int, t_text: int, seed: int = 1, n_images: int | None = None):
    """Synthetic inputs as README.md:34-37 / example.py:5-8 of the reference; n_images = m gives the
    (B, m, 3, H, W) multi-image form of config 5."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(0, cfg.vocab, (batch, t_text), dtype=torch.long, generator=g)
    shape = (batch, 3, cfg.image, cfg.image) if n_images is None else (batch, n_images, 3, cfg.image, cfg.image)
    images = torch.randn(*shape, generator=g)
    return text, images


def build(cfg: OracleConfig | None = None, seed: int = 0, emulate_bf16: bool = False) -> KosmosOracle:
    torch.manual_seed(seed)
    return KosmosOracle(cfg, emulate_bf16).eval()
