"""B200-native drop-in for ``kosmosx.model`` of kyegomez/Kosmos-X.

Same public surface as /root/reference/kosmosx/model.py — ``Kosmos()`` (zero-arg),
``Kosmos.forward(text_tokens, images)``, ``KosmosLanguage``, ``Decoder`` (with
``forward_embedding`` and the ``passed_x`` keyword of README.md:179-193) and the same
``state_dict`` key layout (incl. the multiway ``.A/.B`` branches) — but
every arithmetic step runs in the hand-written sm_100a kernels of libkosmosx_sm100.so through
the C ABI in include/kosmosx_b200.h.  The nn.Modules below only *hold parameters* under the
reference's names; their ``forward`` is never used.  There is no CPU / eager fallback: on a
machine without a B200 the forward raises.

Numerics: bf16 tensor-core operands, fp32 accumulation, fp32 residual stream, fp32
LayerNorm / softmax / GELU (the reference's fp32 ``gelu(x.float())`` and
``softmax(dtype=float32)`` of SURVEY.md A.4).  Dropout (p=0.1) is the eval-mode identity.
"""
from __future__ import annotations

import logging
import math
import os
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import _abi, ops

log = logging.getLogger("kosmosx")


# --------------------------------------------------------------------------- config
@dataclass
class KosmosConfig:
    """Sizes hard-coded at /root/reference/kosmosx/model.py:154-206 (defaults) — keyword-only
    extras of the B200 build let tests shrink the model."""

    vocab: int = 32002
    dim: int = 2048
    layers: int = 24
    ffn: int = 8192
    heads: int = 32
    max_positions: int = 2048
    multiway: bool = True
    xpos_scale_base: int = 512
    eps: float = 1e-5
    vit_dim: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    patch: int = 14
    image: int = 224
    vit_act: str = "gelu"
    p_depth: int = 2
    p_heads: int = 8
    p_dim_head: int = 64
    p_latents: int = 64
    p_media_embeds: int = 257
    p_ff_mult: int = 4
    # torchscale's Decoder.forward_embedding does `x = embed = embed_scale * token_embedding; x += positions` IN PLACE
    # [recall], so the `[1]` result the reference takes at model.py:238 already carries the text positions and the second
    # call (model.py:242-244) adds the spliced positions on top: text rows get TWO positional embeddings.  True reproduces
    # that; False is the out-of-place reading (`x = x + positions`), one positional embedding per row.
    alias_embed_positions: bool = True
    # torchscale DecoderConfig(dropout=0.1, attention_dropout=0.1), reference model.py:175-177: applied by the training
    # step (KosmosTrainer / model.train() forward); Kosmos.forward in eval mode is the identity, as nn.Dropout is
    dropout: float = 0.1
    attention_dropout: float = 0.1

    @property
    def vit_tokens(self) -> int:
        return (self.image // self.patch) ** 2 + 1

    def validate(self):
        if self.dim // self.heads != 64 or self.vit_dim // self.vit_heads != 64 or self.p_dim_head != 64:
            raise ValueError("the sm_100a attention kernels are specialised for head_dim 64")
        for name in ("dim", "ffn", "vit_dim", "vit_mlp"):
            if getattr(self, name) % 64:
                raise ValueError(f"{name} must be a multiple of 64")
        if self.vit_act not in ("gelu", "quick_gelu"):
            raise ValueError("vit_act must be 'gelu' or 'quick_gelu'")
        if self.patch <= 0 or self.image % self.patch or self.image % 4:
            raise ValueError("image must be a multiple of patch and of 4 (the patch pack reads 16-byte pixel vectors)")


# --------------------------------------------------------------------------- parameter containers
class MultiwayNetwork(nn.Module):
    """Container matching torchscale's MultiwayNetwork key layout: ``.A`` is
    live, ``.B`` is kept only so checkpoints round-trip (the reference never runs it)."""

    def __init__(self, make):
        super().__init__()
        self.A = make()
        self.B = make()
        self.split_position = -1


def _mw(cfg: KosmosConfig, make):
    return MultiwayNetwork(make) if cfg.multiway else make()


def _live(m):
    return m.A if isinstance(m, MultiwayNetwork) else m


class _XPosBuffers(nn.Module):
    def __init__(self, head_dim, scale_base):
        super().__init__()
        self.scale_base = scale_base
        self.register_buffer("scale", (torch.arange(0, head_dim, 2) + 0.4 * head_dim) / (1.4 * head_dim))


class _SelfAttnParams(nn.Module):
    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        d = cfg.dim
        self.k_proj = _mw(cfg, lambda: nn.Linear(d, d))
        self.v_proj = _mw(cfg, lambda: nn.Linear(d, d))
        self.q_proj = _mw(cfg, lambda: nn.Linear(d, d))
        self.out_proj = _mw(cfg, lambda: nn.Linear(d, d))
        self.inner_attn_ln = _mw(cfg, lambda: nn.LayerNorm(d, eps=cfg.eps))
        self.xpos = _XPosBuffers(d // cfg.heads, cfg.xpos_scale_base)


class _FFNParams(nn.Module):
    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        self.fc1 = nn.Linear(cfg.dim, cfg.ffn)
        self.fc2 = nn.Linear(cfg.ffn, cfg.dim)
        self.ffn_layernorm = nn.LayerNorm(cfg.ffn, eps=cfg.eps)


class _DecoderLayerParams(nn.Module):
    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        self.self_attn = _SelfAttnParams(cfg)
        self.self_attn_layer_norm = _mw(cfg, lambda: nn.LayerNorm(cfg.dim, eps=cfg.eps))
        self.ffn = _mw(cfg, lambda: _FFNParams(cfg))
        self.final_layer_norm = _mw(cfg, lambda: nn.LayerNorm(cfg.dim, eps=cfg.eps))


class _ClipAttnParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.k_proj, self.v_proj = nn.Linear(d, d), nn.Linear(d, d)
        self.q_proj, self.out_proj = nn.Linear(d, d), nn.Linear(d, d)


class _ClipMLPParams(nn.Module):
    def __init__(self, d, m):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(d, m), nn.Linear(m, d)


class _ClipLayerParams(nn.Module):
    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        self.self_attn = _ClipAttnParams(cfg.vit_dim)
        self.layer_norm1 = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)
        self.mlp = _ClipMLPParams(cfg.vit_dim, cfg.vit_mlp)
        self.layer_norm2 = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)


class _ClipEmbeddingParams(nn.Module):
    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(cfg.vit_dim))
        self.patch_embedding = nn.Conv2d(3, cfg.vit_dim, cfg.patch, cfg.patch, bias=False)
        self.position_embedding = nn.Embedding(cfg.vit_tokens, cfg.vit_dim)


class _ClipEncoderParams(nn.Module):
    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        self.layers = nn.ModuleList(_ClipLayerParams(cfg) for _ in range(cfg.vit_layers))


class ClipVisionTower(nn.Module):
    """Parameter layout of HF ``CLIPVisionTransformer`` ([HF] modeling_clip.py:647-665), i.e. what
    ``CLIPModel.from_pretrained(...).vision_model`` (reference model.py:154-156) exposes.  The
    reference downloads laion/CLIP-ViT-L-14 weights; offline this is random-initialised and the
    real weights arrive through ``load_state_dict``."""

    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        self.embeddings = _ClipEmbeddingParams(cfg)
        self.pre_layrnorm = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)
        self.encoder = _ClipEncoderParams(cfg)
        self.post_layernorm = nn.LayerNorm(cfg.vit_dim, eps=cfg.eps)     # unused on this path


class _PerceiverAttnParams(nn.Module):
    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        d, inner = cfg.vit_dim, cfg.p_heads * cfg.p_dim_head
        self.norm_media = nn.LayerNorm(d)
        self.norm_latents = nn.LayerNorm(d)
        self.to_q = nn.Linear(d, inner, bias=False)
        self.to_kv = nn.Linear(d, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, d, bias=False)


class PerceiverResampler(nn.Module):
    """Parameter layout of flamingo_pytorch.PerceiverResampler."""

    def __init__(self, cfg: KosmosConfig):
        super().__init__()
        d = cfg.vit_dim
        self.latents = nn.Parameter(torch.randn(cfg.p_latents, d))
        self.media_pos_emb = nn.Parameter(torch.randn(cfg.p_media_embeds, 1, d))
        self.layers = nn.ModuleList(
            nn.ModuleList([
                _PerceiverAttnParams(cfg),
                nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d * cfg.p_ff_mult, bias=False), nn.GELU(),
                              nn.Linear(d * cfg.p_ff_mult, d, bias=False)),
            ]) for _ in range(cfg.p_depth))
        self.norm = nn.LayerNorm(d)


# --------------------------------------------------------------------------- engine helpers
def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _bf16(t: torch.Tensor) -> torch.Tensor:
    """fp32 parameter -> bf16 GEMM operand, through kx_cast_f32_to_bf16."""
    return ops.cast_bf16(_f32(t))


def _fold_ln(weight: torch.Tensor, bias, ln: nn.LayerNorm):
    """Fold a LayerNorm into the Linear that consumes it, once, at weight-staging time:
        Linear(LN(x)) = rstd * (x . W'^T - mean * c) + d,   W' = W * gamma,  c = W'.1,  d = W.beta + b
    Returns (W' as bf16, c fp32 [N] summed from the bf16-rounded W' that the tensor cores will see, d fp32 [N])."""
    w = weight.detach().to(torch.float32)
    wb = ops.cast_bf16((w * ln.weight.detach().to(torch.float32)[None, :]).contiguous())
    c = wb.to(torch.float64).sum(dim=1).to(torch.float32).contiguous()
    d = w.to(torch.float64) @ ln.bias.detach().to(torch.float64)
    if bias is not None:
        d = d + bias.detach().to(torch.float64)
    return wb, c, d.to(torch.float32).contiguous()


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: kosmosx (B200 build) has no CPU path — move the module and "
                           "its inputs to a B200 with .cuda()")


class _Workspace:
    """Named activation buffers, allocated once per shape and reused (no allocation in steady state)."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[key] = b
        return b

    def clear(self):
        self._bufs.clear()


# --------------------------------------------------------------------------- incremental decoding state
class DecodeState:
    """KV cache and device-side counters of one generation — what torchscale keeps as
    ``incremental_state[layer]["prev_key" / "prev_value"]``.  Per layer the xPos-rotated keys and the values are bf16
    ``[B, t_max, D]``; ``pos`` (device int32) is the number of cached tokens = the position of the next one, read by the
    kernels themselves so a decoding step needs no host value; ``length`` is its host mirror for argument checks."""

    def __init__(self, cfg: KosmosConfig, batch: int, t_max: int, device):
        if batch > _abi.KX_DECODE_MAX_BATCH:
            raise ValueError(f"incremental decoding handles at most {_abi.KX_DECODE_MAX_BATCH} sequences per call, got {batch}")
        if t_max + 2 > cfg.max_positions:
            raise ValueError(f"prompt + new tokens = {t_max} exceeds the positional table: max is {cfg.max_positions - 2} "
                             "(construct the model with max_positions=... to extend it)")
        bf, f32, i32 = torch.bfloat16, torch.float32, torch.int32
        L, D, F, H = cfg.layers, cfg.dim, cfg.ffn, cfg.heads
        self.batch, self.t_max, self.length = batch, t_max, 0
        self.k = torch.empty(L, batch, H, t_max, D // H, dtype=bf, device=device)      # head-major (kx_decode_attn)
        self.v = torch.empty(L, batch, H, t_max, D // H, dtype=bf, device=device)
        self.keys = torch.zeros(batch, dtype=torch.int64, device=device)               # fused greedy-choice keys
        self.pos = torch.zeros(1, dtype=i32, device=device)
        self.step = torch.zeros(1, dtype=i32, device=device)
        self.counter = torch.zeros(1, dtype=i32, device=device)
        self.err = torch.zeros(1, dtype=i32, device=device)
        self.tok = torch.zeros(batch, dtype=torch.int64, device=device)
        self.x = torch.empty(batch, D, dtype=f32, device=device)
        self.xb = torch.empty(batch, D, dtype=bf, device=device)
        self.q = torch.empty(batch, D, dtype=bf, device=device)
        self.att = torch.empty(batch, D, dtype=bf, device=device)
        self.mid = torch.empty(batch, F, dtype=bf, device=device)
        self.logits = torch.empty(batch, cfg.vocab, dtype=f32, device=device)
        self.scratch, self.counters = ops.decode_attn_scratch(batch, H, t_max, device)
        self.text_off = -1        # >= 0: new tokens also add pos[row + 2 - text_off] (alias_embed_positions; kx_decode_embed)
        self.tabs = None          # xPos tables [4, t_max, 32], centred like the prompt's (set by the prompt pass)
        self.graph = None         # captured (one step + greedy choice), replayed by Kosmos.generate
        self.plan = None          # flattened arguments of the one-kernel step (kx_decode_plan_build)

    def cache_bytes(self) -> int:
        return 2 * self.k.numel() * 2


# --------------------------------------------------------------------------- decoder
class Decoder(nn.Module):
    """torchscale ``Decoder`` surface used by the reference (model.py:186-191,238,242,250) with the
    ``passed_x`` patch of README.md:179-193, executing on the sm_100a kernels."""

    def __init__(self, cfg: KosmosConfig, embed_tokens: nn.Embedding, embed_positions: nn.Embedding,
                 output_projection: nn.Linear):
        super().__init__()
        self.cfg = cfg
        self.embed_scale = 1.0
        self.embed_tokens = embed_tokens
        self.embed_positions = embed_positions
        self.output_projection = output_projection
        self.layers = nn.ModuleList(_DecoderLayerParams(cfg) for _ in range(cfg.layers))
        self.layer_norm = nn.LayerNorm(cfg.dim, eps=cfg.eps)
        init_scale = math.sqrt(math.log(cfg.layers * 2))              # sub-LN init
        for name, p in self.named_parameters():
            if "fc1" in name or "fc2" in name or "out_proj" in name or "v_proj" in name:
                p.data.mul_(init_scale)
        self._packed = None
        self._xpos_cache = {}
        self._ws = _Workspace()

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.invalidate()
        return out

    # ---- weight staging ---------------------------------------------------------------
    def invalidate(self):
        self._packed = None
        self._xpos_cache.clear()
        self._ws.clear()

    def _pack(self):
        if self._packed is not None:
            return self._packed
        _require_cuda(self.layer_norm.weight, "Decoder parameters")
        layers = []
        for L in self.layers:
            sa = L.self_attn
            q, k, v, o = (_live(m) for m in (sa.q_proj, sa.k_proj, sa.v_proj, sa.out_proj))
            ffn = _live(L.ffn)
            ln_a, ln_i, ln_f = _live(L.self_attn_layer_norm), _live(sa.inner_attn_ln), _live(L.final_layer_norm)
            # every LayerNorm of the layer is folded into the Linear that consumes it
            layers.append(dict(
                qkv=_fold_ln(torch.cat([q.weight, k.weight, v.weight], 0), torch.cat([q.bias, k.bias, v.bias], 0), ln_a),
                o=_fold_ln(o.weight, o.bias, ln_i),
                fc1=_fold_ln(ffn.fc1.weight, ffn.fc1.bias, ln_f),
                fc2=_fold_ln(ffn.fc2.weight, ffn.fc2.bias, ffn.ffn_layernorm),
            ))
        self._packed = dict(
            layers=layers,
            out=_fold_ln(self.output_projection.weight, self.output_projection.bias, self.layer_norm),
            embed=_f32(self.embed_tokens.weight), pos=_f32(self.embed_positions.weight),
            xpos_scale=_f32(self.layers[0].self_attn.xpos.scale) if len(self.layers) else None,
        )
        return self._packed

    def _xpos(self, T: int, device):
        tabs = self._xpos_cache.get(T)
        if tabs is None:
            hd = self.cfg.dim // self.cfg.heads
            half = hd // 2
            # inv_freq exactly as the reference computes it on the host
            inv_freq = (1.0 / (10000 ** (torch.arange(0, half) / half))).to(device=device, dtype=torch.float32)
            scale = _f32(self.layers[0].self_attn.xpos.scale)     # the buffer itself: no weight staging from here
            tabs = ops.xpos_tables(scale, inv_freq, T, (-T) // 2, float(self.cfg.xpos_scale_base), device)
            self._xpos_cache[T] = tabs
        return tabs

    # ---- reference API ----------------------------------------------------------------
    def forward_embedding(self, tokens, token_embedding=None, incremental_state=None):
        """(x, embed) as torchscale's Decoder.forward_embedding.  ``tokens`` may be a
        float (B,T,D) tensor when ``token_embedding`` is given — only its length is used.  With
        ``cfg.alias_embed_positions`` (torchscale's in-place ``x += positions`` on ``x = embed = ...``) the second result
        IS the first: ``embed`` carries the positions too, which is what model.py:238 then splices."""
        p = self._pack()
        T = tokens.size(1)
        if T + 2 > p["pos"].shape[0]:
            raise ValueError(f"sequence length {T} exceeds the positional table ({p['pos'].shape[0]} rows, max T = "
                             f"{p['pos'].shape[0] - 2})")
        if token_embedding is None:
            _require_cuda(tokens, "tokens")
            self._check_tokens(tokens)
            embed = torch.empty(tokens.shape[0], T, self.cfg.dim, dtype=torch.float32, device=tokens.device)
            ops.embed_splice_pos(tokens, p["embed"], None, embed)
        else:
            _require_cuda(token_embedding, "token_embedding")
            embed = token_embedding.to(torch.float32).contiguous()
        x = torch.empty_like(embed)
        ops.add_positions(embed, x, p["pos"])
        return (x, x) if self.cfg.alias_embed_positions else (x, embed)

    def _check_tokens(self, tokens):
        if tokens.dtype != torch.int64:
            raise TypeError("token ids must be int64")

    def run_layers(self, x: torch.Tensor, B: int, T: int, logits: torch.Tensor | None = None, head: bool = True,
                   state: DecodeState | None = None, keep: dict | None = None):
        """24 x (sub-LN attention + sub-LN FFN) in place on the fp32 residual stream x [B*T, D], then (``head``)
        final LayerNorm + LM head -> fp32 logits [B*T, vocab].  5 launches per layer: every LayerNorm is folded
        into its consumer GEMM (row statistics travel as partial sums from the producer's epilogue)."""
        cfg, p, ws = self.cfg, self._pack(), self._ws
        M, D, F, H = B * T, cfg.dim, cfg.ffn, cfg.heads
        dev = x.device
        bf, f32 = torch.bfloat16, torch.float32
        xb = ws.get("xb", (M, D), bf, dev)                       # bf16 copy of the residual stream (GEMM operand)
        qkv = ws.get("qkv", (M, 3 * D), bf, dev)
        att = ws.get("att", (M, D), bf, dev)
        mid = ws.get("mid", (M, F), bf, dev)
        # per-row partial (sum, sumsq) emitted by each producer, summed by the consumer's folded LayerNorm
        st_in = ws.get("st_in", (1, M, 2), f32, dev)
        st_a = ws.get("st_a", ((D + 127) // 128, M, 2), f32, dev)
        st_b = ws.get("st_b", ((D + 127) // 128, M, 2), f32, dev)
        st_att = ws.get("st_att", (H, M, 2), f32, dev)
        st_mid = ws.get("st_mid", ((F + 127) // 128, M, 2), f32, dev)
        # prompt pass of a generation (``state``): the tables extend to t_max rows with the prompt's centre, so rows
        # 0..T-1 are the ordinary ones, and every layer's rotated k / v go into the cache
        tabs = self._xpos(T, dev) if state is None else state.tabs
        scale = (D // H) ** -0.5
        eps = cfg.eps
        ops.rowstats_cast(x, xb, st_in)
        cur = st_in
        for li, L in enumerate(p["layers"]):
            w, c, d = L["qkv"]                                   # self_attn_layer_norm -> q|k|v (+bias, xPos)
            ops.gemm(xb, w, qkv, bias=d, ln=(cur, c, D, eps), xpos=(tabs[0], tabs[1], tabs[2], tabs[3]), seq_len=T)
            if state is not None:
                ops.kv_cache_store(qkv, state.k[li], state.v[li], batch=B, seq_len=T, d_model=D, t_max=state.t_max)
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, batch=B, heads=H, seq_len=T, causal=True,
                          scale=scale, stats_out=st_att)
            w, c, d = L["o"]                                     # inner_attn_ln -> out_proj (+bias, +residual)
            ops.gemm(att, w, x, bias=d, res=x, ln=(st_att, c, D, eps), stats_out=st_a, out2=xb)
            w, c, d = L["fc1"]                                   # final_layer_norm -> fc1 (+bias, GELU)
            ops.gemm(xb, w, mid, bias=d, act=_abi.KX_ACT_GELU, ln=(st_a, c, D, eps), stats_out=st_mid)
            w, c, d = L["fc2"]                                   # ffn_layernorm -> fc2 (+bias, +residual)
            ops.gemm(mid, w, x, bias=d, res=x, ln=(st_mid, c, F, eps), stats_out=st_b, out2=xb)
            cur = st_b
            if keep is not None:                                 # per-stage parity tables (tools/stage_errors.py)
                keep.setdefault("inner_states", []).append(x.view(B, T, D).clone())
        if not head:
            return x
        w, c, d = p["out"]                                       # decoder.layer_norm -> output_projection
        if logits is None:
            logits = torch.empty(M, w.shape[0], dtype=torch.float32, device=dev)
        ops.gemm(xb, w, logits[:, :w.shape[0]], bias=d, ln=(cur, c, D, eps))      # bf16 logits: rows padded to 16 bytes
        return logits

    # ---- incremental decoding ------------------------------------------
    def begin_generation(self, x: torch.Tensor, B: int, T: int, t_max: int, head="all", text_off: int = -1) -> tuple[DecodeState, torch.Tensor]:
        """Prompt pass: the ordinary layers over x [B*T, D] (updated in place) with every layer's rotated k / v stored
        in a fresh cache of t_max rows.  head = "all": logits of every row [B*T, vocab]; "last": only the last row of
        each sequence, in state.logits [B, vocab] (what a generation loop needs)."""
        if t_max < T:
            raise ValueError(f"t_max {t_max} is shorter than the prompt ({T})")
        cfg, p = self.cfg, self._pack()
        state = DecodeState(cfg, B, t_max, x.device)
        state.text_off = text_off
        half = cfg.dim // cfg.heads // 2
        inv_freq = (1.0 / (10000 ** (torch.arange(0, half) / half))).to(device=x.device, dtype=torch.float32)
        state.tabs = ops.xpos_tables(p["xpos_scale"], inv_freq, t_max, (-T) // 2, float(cfg.xpos_scale_base), x.device)
        if head == "all":
            out = self.run_layers(x, B, T, state=state)
        else:
            self.run_layers(x, B, T, head=False, state=state)
            xb = self._ws.get("xb", (B * T, cfg.dim), torch.bfloat16, x.device)
            w, c, d = p["out"]                               # decoder.layer_norm -> output_projection, last rows only
            ops.decode_linear(xb.view(B, T, cfg.dim)[:, T - 1], w, bias=d, ln_c=c, eps=cfg.eps, out=state.logits,
                              argmax_keys=state.keys)
            out = state.logits
        state.pos.fill_(T)
        state.length = T
        return state, out

    def decode_step(self, state: DecodeState):
        """One new token per sequence: embeds state.tok at position state.pos, appends its k / v to the cache, leaves
        the next-token logits in state.logits.  Does not advance the position (kx_argmax_advance does).  5 launches
        per layer, all reading the position from device memory, so the sequence is CUDA-graph replayable."""
        p = self._pack()
        ops.decode_embed(state.tok, p["embed"], p["pos"], state.pos, state.x, state.xb, state.err, text_index_off=state.text_off)
        self._decode_layers(state)

    def _decode_layers(self, state: DecodeState):
        cfg, p = self.cfg, self._pack()
        D, H, eps = cfg.dim, cfg.heads, cfg.eps
        scale = (D // H) ** -0.5
        for li, L in enumerate(p["layers"]):
            w, c, d = L["qkv"]
            ops.decode_linear(state.xb, w, bias=d, ln_c=c, eps=eps,
                              qkv=(state.q, state.k[li], state.v[li], state.t_max, state.pos, state.tabs))
            ops.decode_attention(state.q, state.k[li], state.v[li], state.att, t_max=state.t_max, heads=H, pos=state.pos,
                                 scale=scale, scratch=state.scratch, counters=state.counters)
            w, c, d = L["o"]
            ops.decode_linear(state.att, w, bias=d, ln_c=c, eps=eps, res=(state.x, state.xb))
            w, c, d = L["fc1"]
            ops.decode_linear(state.xb, w, bias=d, ln_c=c, eps=eps, act=_abi.KX_ACT_GELU, out=state.mid)
            w, c, d = L["fc2"]
            ops.decode_linear(state.mid, w, bias=d, ln_c=c, eps=eps, res=(state.x, state.xb))
        w, c, d = p["out"]                                   # + the greedy choice, reduced into state.keys
        ops.decode_linear(state.xb, w, bias=d, ln_c=c, eps=eps, out=state.logits, argmax_keys=state.keys)

    def build_step_plan(self, state: DecodeState, history=None, forced=None, trace=None):
        """Flatten one generation's arguments for the persistent one-kernel step (kx_decode_step; batch <= 8)."""
        cfg, p = self.cfg, self._pack()
        plan, barrier = ops.decode_step_buffers(cfg.layers, state.x.device)
        state.step_bufs = (barrier,)
        state.plan = ops.decode_plan_build(
            plan, layers=p["layers"], out=p["out"], embed_table=p["embed"], pos_table=p["pos"], tabs=state.tabs,
            k_cache=[state.k[i] for i in range(cfg.layers)], v_cache=[state.v[i] for i in range(cfg.layers)],
            tokens=state.tok, x=state.x, xb=state.xb, q=state.q, att=state.att, mid=state.mid, logits=state.logits,
            keys=state.keys, pos=state.pos, step=state.step, err_flag=state.err, barrier=barrier, heads=cfg.heads, ffn=cfg.ffn, t_max=state.t_max, eps=cfg.eps,
            scale=(cfg.dim // cfg.heads) ** -0.5, forced=forced, history=history, trace=trace, text_index_off=state.text_off)
        return state.plan

    def advance(self, state: DecodeState, history=None, forced=None, move=True):
        """Greedy choice (reduced by the LM-head launch into state.keys) or the forced token into state.tok, optional
        history column, position += 1."""
        ops.argmax_advance(state.logits, state.tok, step=state.step, counter=state.counter,
                           pos=state.pos if move else None, history=history, forced=forced, keys=state.keys)
        if move:
            state.length += 1

    def forward(self, prev_output_tokens, incremental_state=None, token_embeddings=None, **kwargs):
        """Returns ``(logits, extra)``; ``passed_x`` (B,T,D) skips the embedding (README.md:179-193).

        ``incremental_state`` follows torchscale's generation protocol (SURVEY.md A.4 [recall]): create it as
        ``{"is_first_step": True}`` for the prompt pass (all rows run and are cached; optional ``"max_length"`` sizes the
        cache, default = the positional table), set ``is_first_step`` to ``False`` and call again with the whole prefix
        (only its last token is embedded, at position ``len(prefix) + 1``) for every following token; logits are then
        (B, 1, vocab).  The cache lives under ``incremental_state["kx_state"]`` (a DecodeState)."""
        passed = kwargs.get("passed_x", None)
        extra = {"inner_states": None, "l_aux": [None] * len(self.layers), "attn": None}
        if incremental_state is not None and not incremental_state.get("is_first_step", False):
            state = incremental_state.get("kx_state")
            if state is None:
                raise ValueError("incremental_state has no cache: run the prompt pass with is_first_step=True first")
            if state.length + 1 > state.t_max:
                raise ValueError(f"the KV cache is full ({state.t_max} rows); pass a larger incremental_state['max_length']")
            if passed is not None:                                          # an already embedded (B,1,D) row
                _require_cuda(passed, "passed_x")
                if passed.shape[0] != state.batch or passed.shape[1] != 1:
                    raise ValueError("after the first step passed_x must be (B, 1, D)")
                state.x.copy_(passed[:, 0])
                ops.cast_bf16(state.x, state.xb)
                self._decode_layers(state)
            else:
                _require_cuda(prev_output_tokens, "prev_output_tokens")
                self._check_tokens(prev_output_tokens)
                if prev_output_tokens.shape[0] != state.batch or prev_output_tokens.shape[1] != state.length + 1:
                    raise ValueError(f"prev_output_tokens must be the whole prefix, (B={state.batch}, {state.length + 1}); got "
                                     f"{tuple(prev_output_tokens.shape)}")
                state.tok.copy_(prev_output_tokens[:, -1])
                self.decode_step(state)
            logits = state.logits.clone()
            self.advance(state)
            return logits.view(state.batch, 1, -1), extra
        if passed is None:
            x, _ = self.forward_embedding(prev_output_tokens, token_embeddings)
        else:
            _require_cuda(passed, "passed_x")
            x = passed.to(torch.float32).clone()          # layers update the stream in place
        B, T, D = x.shape
        if incremental_state is not None:                 # prompt pass of a generation
            t_max = int(incremental_state.get("max_length", self.cfg.max_positions - 2))
            state, logits = self.begin_generation(x.view(B * T, D), B, T, t_max)
            incremental_state["kx_state"] = state
            return logits.view(B, T, -1), extra
        logits = self.run_layers(x.view(B * T, D), B, T).view(B, T, -1)
        return logits, extra


# --------------------------------------------------------------------------- shared top-level plumbing
class _KosmosBase(nn.Module):
    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.refresh_weights()
        return out

    def refresh_weights(self):
        """Drop the staged bf16 weights / workspaces; they are rebuilt on the next forward.  Called
        automatically after ``.to()/.cuda()`` and ``load_state_dict``; call it yourself after
        mutating parameters in place."""
        if hasattr(self, "decoder"):
            self.decoder.invalidate()
        self._vis_packed = None
        if hasattr(self, "_ws"):
            self._ws.clear()
        self._graphs = {}
        if getattr(self, "_acc", None) is not None:
            self._acc.invalidate()

    _PRECISIONS = ("bf16", "bf16x3")

    def _accurate(self):
        """The verification-precision engine (kosmosx/accurate.py), built on first use."""
        if getattr(self, "_acc", None) is None:
            from .accurate import AccurateEngine
            self._acc = AccurateEngine(self)
        return self._acc

    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        self.refresh_weights()
        return r

    # ---- checkpoint interchange ------------------------------------------
    _WRAPPER_PREFIXES = ("module.", "_fsdp_wrapped_module.", "_orig_mod.")     # DDP / FSDP / torch.compile wrappers
    _TIED = (("embed.weight", "decoder.embed_tokens.weight"),                  # Appendix B: same tensor, both keys saved
             ("embed_positions.weight", "decoder.embed_positions.weight"),
             ("output_projection.weight", "decoder.output_projection.weight"))

    def save_checkpoint(self, path):
        """The reference's format (train.py:688-695): ``torch.save(unwrapped_model.state_dict(), final_model.pt)``."""
        torch.save(self.state_dict(), path)

    def load_checkpoint(self, checkpoint, strict: bool = True, resize_positions: bool = False):
        """Load a reference ``final_model.pt`` (train.py:688-695) or an already loaded state_dict.

        Beyond ``load_state_dict``: wrapper prefixes left by DDP / FSDP / torch.compile are stripped; HF's legacy
        ``embeddings.position_ids`` buffer is dropped; a tied pair saved under only one of its two names (Appendix B)
        is completed; a plain (non-multiway) torchscale decoder checkpoint is mapped onto the live ``.A`` branches;
        dtypes are converted by ``copy_``.  ``resize_positions`` copies the overlapping rows of a positional table of
        another length (the reference's has 2048 rows) instead of failing.  Returns the
        ``(missing_keys, unexpected_keys)`` result of ``load_state_dict``."""
        if not isinstance(checkpoint, dict):
            checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=True)
        if "state_dict" in checkpoint and isinstance(checkpoint["state_dict"], dict):
            checkpoint = checkpoint["state_dict"]
        own = self.state_dict()
        sd = {}
        for k, v in checkpoint.items():
            for w in self._WRAPPER_PREFIXES[1:]:
                k = k.replace(w, "")
            while k.startswith("module."):
                k = k[len("module."):]
            if k.endswith("embeddings.position_ids"):
                continue
            sd[k] = v
        for a, b in self._TIED:
            if a in own and b in own:
                if a in sd and b not in sd:
                    sd[b] = sd[a]
                elif b in sd and a not in sd:
                    sd[a] = sd[b]
        def live_branch(k):                              # non-multiway torchscale name -> the live .A branch's name
            head, _, leaf = k.rpartition(".")
            parent, _, mod = head.rpartition(".")
            for cand in (f"{head}.A.{leaf}",             # self_attn.q_proj.weight -> self_attn.q_proj.A.weight
                         f"{parent}.A.{mod}.{leaf}"):    # ffn.fc1.weight -> ffn.A.fc1.weight
                if cand in own and cand not in sd:
                    return cand
            return None

        plain = {k: live_branch(k) for k in sd if k not in own and k.startswith("decoder.layers.")}
        plain = {k: c for k, c in plain.items() if c is not None}
        for k, cand in plain.items():
            sd[cand] = sd.pop(k)
        if resize_positions:
            for k in ("embed_positions.weight", "decoder.embed_positions.weight"):
                if k in sd and k in own and sd[k].shape != own[k].shape and sd[k].shape[1:] == own[k].shape[1:]:
                    rows = min(sd[k].shape[0], own[k].shape[0])
                    merged = own[k].detach().clone()
                    merged[:rows] = sd[k][:rows].to(merged.dtype)
                    sd[k] = merged
        res = self.load_state_dict(sd, strict=strict and not plain)
        if plain and strict:                             # the .B branches (never run, a17) may keep their initial values; nothing else
            missing = [k for k in res.missing_keys if ".B." not in k]
            if missing or res.unexpected_keys:
                raise RuntimeError(f"load_checkpoint: missing keys {missing}, unexpected keys {list(res.unexpected_keys)}")
        return res

    def _generate(self, x0: torch.Tensor, B: int, T: int, max_new_tokens: int, forced=None, return_logits=False,
                  cuda_graph=True, one_kernel=None, text_off: int = -1):
        """Greedy continuation of the embedded prompt x0 [B*T, D]: prompt pass with cache fill and
        the LM head on the last rows only, then max_new_tokens - 1 one-token steps.  The step + greedy choice is
        captured once as a CUDA graph and replayed (the position is device-resident), so no host value is read until
        the token history is returned.  ``forced`` (B, n) int64 replaces the greedy choice (teacher forcing);
        ``return_logits`` also returns the (B, n, vocab) distributions (eager steps, one clone per step)."""
        if max_new_tokens < 1:
            raise ValueError("max_new_tokens must be >= 1")
        dec = self.decoder
        dev = x0.device
        state, _ = dec.begin_generation(x0, B, T, T + max_new_tokens, head="last", text_off=text_off)
        history = torch.zeros(B, max_new_tokens, dtype=torch.int64, device=dev)
        if forced is not None:
            _require_cuda(forced, "forced tokens")
            if forced.dtype != torch.int64 or tuple(forced.shape) != (B, max_new_tokens):
                raise TypeError(f"forced tokens must be int64 (B={B}, {max_new_tokens})")
            forced = forced.contiguous()
        outs = [state.logits.clone()] if return_logits else None
        dec.advance(state, history=history, forced=forced, move=False)       # token 0 comes from the prompt's last row
        steps = max_new_tokens - 1
        if one_kernel is None:
            one_kernel = B <= 8
            auto = True
        else:
            auto = False
        plan = None
        if one_kernel and steps > 0:
            # the whole step (embedding .. greedy choice) is ONE persistent cooperative launch per token
            if B > 8:
                raise ValueError("the one-kernel decoding step handles at most 8 sequences")
            try:
                plan = dec.build_step_plan(state, history=history, forced=forced)
            except RuntimeError:
                if not auto:                                 # asked for explicitly: report it
                    raise
                log.warning("one-kernel decoding step unavailable (cooperative launch does not fit); using the per-kernel step")
        if plan is not None:
            for _ in range(steps):
                ops.decode_step(plan)
                state.length += 1
                if return_logits:
                    outs.append(state.logits.clone())
        elif return_logits or not cuda_graph or steps < 2:
            for _ in range(steps):
                dec.decode_step(state)
                if return_logits:
                    outs.append(state.logits.clone())
                dec.advance(state, history=history, forced=forced)
        else:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(graph):                # capture only records: nothing runs, the cache is untouched
                dec.decode_step(state)
                dec.advance(state, history=history, forced=forced)
            nodes = ops.launch_count() - n0
            state.length -= 1                            # the captured call counted a step that has not run yet
            state.graph = graph
            for _ in range(steps):
                graph.replay()
                ops.count_graph_replay(nodes)
                state.length += 1
        self._last_decode_state = state
        if return_logits:
            return history, torch.stack(outs, 1)
        return history


def _embed_init(emb: nn.Embedding):
    nn.init.xavier_uniform_(emb.weight)          # bitsandbytes.nn.Embedding.reset_parameters
    with torch.no_grad():
        emb.weight[emb.padding_idx].fill_(0)


class Kosmos(_KosmosBase):
    """``kosmosx.model.Kosmos`` (reference model.py:132-253): CLIP ViT-L/14 -> PerceiverResampler
    (257 -> 64 latents) -> image_proj -> splice after token 1 -> 24-layer sub-LN / xPos decoder.

    ``Kosmos()`` takes no positional arguments, like the reference.  Keyword-only extras:
    ``config`` (KosmosConfig), ``device``, ``max_positions`` (the reference's 2048-row table caps the
    spliced length at 2046), ``cuda_graph`` (replay the whole forward as one CUDA
    graph per input shape; the logits are copied out of the graph's buffer so the caller owns them, as with the
    reference — ``graph_alias_output=True`` skips that copy and returns a view of one of two alternating graph
    buffers, valid until the second-next call with the same shapes), ``logits_dtype`` (torch.float32, or
    torch.bfloat16 = what the reference returns when the module is cast to bf16: the LM head stores bf16 rows padded
    to a 16-byte pitch through the TMA-store epilogue and ``forward`` returns the (B, T, vocab) view of them),
    ``precision``: "bf16" (tensor-core operands rounded to bf16, the speed mode) or "bf16x3" (verification
    mode: split-operand GEMMs + fp32 everywhere else, logits within 1e-3 of the fp32 reference —
    kosmosx/accurate.py; may also be switched per call with ``forward(..., precision=...)``).
    """

    def __init__(self, *, config: KosmosConfig | None = None, device=None, max_positions: int | None = None,
                 cuda_graph: bool = False, precision: str = "bf16", graph_alias_output: bool = False,
                 logits_dtype: torch.dtype = torch.float32):
        super().__init__()
        if logits_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("logits_dtype must be torch.float32 or torch.bfloat16")
        self.logits_dtype = logits_dtype
        self.graph_alias_output = graph_alias_output
        if precision not in self._PRECISIONS:
            raise ValueError(f"precision must be one of {self._PRECISIONS}")
        self.precision = precision
        cfg = config or KosmosConfig()
        if max_positions is not None:
            cfg.max_positions = max_positions
        cfg.validate()
        self.cfg = cfg
        self.cuda_graph = cuda_graph
        self._vis_packed = None
        self._ws = _Workspace()
        self._graphs = {}
        with torch.device(device) if device is not None else _nullctx():
            self.clip_model = ClipVisionTower(cfg)                                   # model.py:154-156
            self.embed = nn.Embedding(cfg.vocab, cfg.dim, padding_idx=1)             # model.py:161-163
            _embed_init(self.embed)
            self.embed_positions = nn.Embedding(cfg.max_positions, cfg.dim, padding_idx=1)   # model.py:164
            self.output_projection = nn.Linear(cfg.dim, cfg.vocab, bias=False)       # model.py:166-167
            nn.init.normal_(self.output_projection.weight, mean=0, std=cfg.dim ** -0.5)
            self.config = cfg                                                        # model.py:170 (DecoderConfig)
            self.decoder = Decoder(cfg, self.embed, self.embed_positions, self.output_projection)
            self.perceive = PerceiverResampler(cfg)                                  # model.py:196-203
            self.image_proj = nn.Linear(cfg.vit_dim, cfg.dim, bias=False)            # model.py:205-206
            nn.init.normal_(self.image_proj.weight, mean=0, std=cfg.dim ** -0.5)
        self.eval()

    # ---- staging ------------------------------------------------------------------------
    def _pack_resampler(self):
        """Staged weights of the perceiver resampler + image_proj (re-derived alone after a training step)."""
        cfg = self.cfg
        pl = []
        for attn, ff in self.perceive.layers:
            pl.append(dict(
                nm=(_f32(attn.norm_media.weight), _f32(attn.norm_media.bias)),
                nl=(_f32(attn.norm_latents.weight), _f32(attn.norm_latents.bias)),
                w_q=_bf16(attn.to_q.weight), w_kv=_bf16(attn.to_kv.weight), w_out=_bf16(attn.to_out.weight),
                ff_ln=(_f32(ff[0].weight), _f32(ff[0].bias)), w_ff1=_bf16(ff[1].weight), w_ff2=_bf16(ff[3].weight),
            ))
        return dict(p_layers=pl, latents=_f32(self.perceive.latents),
                    media_pos=_f32(self.perceive.media_pos_emb).view(-1, cfg.vit_dim),
                    p_norm=(_f32(self.perceive.norm.weight), _f32(self.perceive.norm.bias)), w_ip=_bf16(self.image_proj.weight))

    @staticmethod
    def _pack_clip_layer(L):
        """One CLIP encoder layer staged for the 5-launch inference layer: layer_norm1 / layer_norm2 are folded into q|k|v and
        fc1 (same scheme as the decoder)."""
        a = L.self_attn
        return dict(
            qkv=_fold_ln(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0),
                         torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0), L.layer_norm1),
            w_o=_bf16(a.out_proj.weight), b_o=_f32(a.out_proj.bias),
            fc1=_fold_ln(L.mlp.fc1.weight, L.mlp.fc1.bias, L.layer_norm2),
            w_fc2=_bf16(L.mlp.fc2.weight), b_fc2=_f32(L.mlp.fc2.bias),
        )

    def _pack_vision(self):
        if self._vis_packed is not None:
            if getattr(self, "_resampler_dirty", False):
                self._vis_packed.update(self._pack_resampler())
                self._resampler_dirty = False
            if getattr(self, "_clip_last_dirty", False):          # KosmosTrainer(train_clip_last_layer=True) stepped it
                self._vis_packed["layers"][-1] = self._pack_clip_layer(self.clip_model.encoder.layers[-1])
                self._clip_last_dirty = False
            return self._vis_packed
        cfg, cm = self.cfg, self.clip_model
        _require_cuda(cm.pre_layrnorm.weight, "Kosmos parameters")
        k = 3 * cfg.patch * cfg.patch
        k_pad = (k + 63) // 64 * 64
        wp = torch.zeros(cfg.vit_dim, k_pad, dtype=torch.float32, device=cm.pre_layrnorm.weight.device)
        wp[:, :k] = cm.embeddings.patch_embedding.weight.detach().reshape(cfg.vit_dim, k)
        layers = [self._pack_clip_layer(L) for L in cm.encoder.layers]
        self._clip_last_dirty = False
        self._vis_packed = dict(
            k_pad=k_pad, w_patch=_bf16(wp), cls=_f32(cm.embeddings.class_embedding),
            vpos=_f32(cm.embeddings.position_embedding.weight),
            pre_ln=(_f32(cm.pre_layrnorm.weight), _f32(cm.pre_layrnorm.bias)),
            layers=layers,
        )
        self._vis_packed.update(self._pack_resampler())
        self._resampler_dirty = False
        return self._vis_packed

    # ---- stages -------------------------------------------------------------------------
    def _vit(self, images: torch.Tensor, media: int = 1, upto: int | None = None) -> torch.Tensor:
        """CLIPVisionTransformer.forward ([HF] modeling_clip.py:667-697) -> fp32 [N*Tv, Dv] (un-normalised) for the
        N = images.shape[0] images.  media > 1: `images` is (sequences*media, 3, H, W) in (sequence, media) order and
        the output rows are media-major (image i of every sequence is one contiguous block).  uint8 `images`
        ((N,3,H,W) or (N,H,W,3)) are raw pixels: CLIP's rescale + normalise is fused into the patch pack.
        upto: run only encoder layers [0, upto) (KosmosTrainer fine-tuning the last layer runs that one itself)."""
        cfg, vp, ws = self.cfg, self._pack_vision(), self._ws
        B = images.shape[0]
        Tv, Dv, P = cfg.vit_tokens, cfg.vit_dim, cfg.vit_tokens - 1
        M = B * Tv
        dev = images.device
        patches = ws.get("patches", (B * P, vp["k_pad"]), torch.bfloat16, dev)
        emb = ws.get("vemb", (M, Dv), torch.float32, dev)
        x = ws.get("vx", (M, Dv), torch.float32, dev)
        qkv = ws.get("vqkv", (M, 3 * Dv), torch.bfloat16, dev)
        att = ws.get("vatt", (M, Dv), torch.bfloat16, dev)
        mid = ws.get("vmid", (M, cfg.vit_mlp), torch.bfloat16, dev)
        pack = ops.im2col_patches_u8 if images.dtype == torch.uint8 else ops.im2col_patches   # uint8: raw pixels (§8(f)4)
        pack(images, patches, vp["cls"], vp["vpos"], emb.view(B, Tv, Dv), image=cfg.image, patch=cfg.patch, media=media)
        ops.gemm(patches, vp["w_patch"], emb, grp=(P, Tv, 1), add_tab=vp["vpos"], add_off=1)
        ops.layernorm(emb, *vp["pre_ln"], x, eps=cfg.eps)                    # fp32 out: the residual stream
        act = _abi.KX_ACT_GELU if cfg.vit_act == "gelu" else _abi.KX_ACT_QUICK_GELU
        scale = (Dv // cfg.vit_heads) ** -0.5
        xb = ws.get("vxb", (M, Dv), torch.bfloat16, dev)                      # bf16 copy of the stream (GEMM operand)
        st0 = ws.get("vst0", (1, M, 2), torch.float32, dev)
        # 64-column partials = 128-wide tiles for the N = Dv GEMMs: twice as many tiles to spread over the SMs
        st_a = ws.get("vst_a", ((Dv + 63) // 64, M, 2), torch.float32, dev)
        st_b = ws.get("vst_b", ((Dv + 63) // 64, M, 2), torch.float32, dev)
        ops.rowstats_cast(x, xb, st0)
        cur = st0
        for L in vp["layers"][:upto]:                                        # 5 launches per layer, no stand-alone LayerNorm
            w, c, d = L["qkv"]
            ops.gemm(xb, w, qkv, bias=d, ln=(cur, c, Dv, cfg.eps))
            ops.attention(qkv[:, :Dv], qkv[:, Dv:2 * Dv], qkv[:, 2 * Dv:], att, batch=B, heads=cfg.vit_heads,
                          seq_len=Tv, causal=False, scale=scale)
            ops.gemm(att, L["w_o"], x, bias=L["b_o"], res=x, stats_out=st_a, out2=xb)
            w, c, d = L["fc1"]
            ops.gemm(xb, w, mid, bias=d, act=act, ln=(st_a, c, Dv, cfg.eps))
            ops.gemm(mid, L["w_fc2"], x, bias=L["b_fc2"], res=x, stats_out=st_b, out2=xb)
            cur = st_b
        return x

    def _perceive_project(self, xv: torch.Tensor, B: int, x0: torch.Tensor, T: int, img_rows=(2,), pos_table=None):
        """PerceiverResampler + image_proj (model.py:232).  xv holds m = len(img_rows) images per
        sequence, media-major ([m*B*Tv, Dv]); the projection's epilogue writes rows [img_rows[i], img_rows[i]+64) of
        every sequence of x0 and adds their positions."""
        cfg, vp, ws = self.cfg, self._pack_vision(), self._ws
        Tv, Dv, Lq, Hp = cfg.vit_tokens, cfg.vit_dim, cfg.p_latents, cfg.p_heads
        inner = Hp * cfg.p_dim_head
        dev = xv.device
        m = len(img_rows)
        if m > cfg.p_media_embeds:
            raise ValueError(f"{m} images per sequence but media_pos_emb has {cfg.p_media_embeds} rows")
        N = B * m                                        # every (image, sequence) pair resamples independently (A.2)
        lat = ws.get("plat", (N * Lq, Dv), torch.float32, dev)
        cat = ws.get("pcat", (N * (Tv + Lq), Dv), torch.bfloat16, dev)
        lnl = ws.get("plnl", (N * Lq, Dv), torch.bfloat16, dev)
        q = ws.get("pq", (N * Lq, inner), torch.bfloat16, dev)
        kv = ws.get("pkv", (N * (Tv + Lq), 2 * inner), torch.bfloat16, dev)
        att = ws.get("patt", (N * Lq, inner), torch.bfloat16, dev)
        mid = ws.get("pmid", (N * Lq, Dv * cfg.p_ff_mult), torch.bfloat16, dev)
        ops.broadcast_rows(vp["latents"], lat, N)
        mp = vp["media_pos"][0:m]                        # media_pos_emb[:m]: row i goes to image i (media-major blocks)
        for L in vp["p_layers"]:
            ops.layernorm(xv, *L["nm"], cat, pre_add=mp, pre_add_group=Tv * B if m > 1 else 0, grp=(Tv, Tv + Lq, 0))
            ops.layernorm(lat, *L["nl"], cat, grp=(Lq, Tv + Lq, Tv))
            ops.layernorm(lat, *L["nl"], lnl)
            ops.gemm(lnl, L["w_q"], q)
            ops.gemm(cat, L["w_kv"], kv)
            ops.perceiver_attention(q, kv, att, batch=N, heads=Hp, n_q=Lq, n_kv=Tv + Lq, v_col_off=inner,
                                    scale=cfg.p_dim_head ** -0.5)
            ops.gemm(att, L["w_out"], lat, res=lat)
            ops.layernorm(lat, *L["ff_ln"], lnl)
            ops.gemm(lnl, L["w_ff1"], mid, act=_abi.KX_ACT_GELU)
            ops.gemm(mid, L["w_ff2"], lat, res=lat)
        ops.layernorm(lat, *vp["p_norm"], lnl)
        pos = pos_table if pos_table is not None else self.decoder._pack()["pos"]
        for i, r0 in enumerate(img_rows):                # image i of all sequences: rows [i*B*64, (i+1)*B*64)
            ops.gemm(lnl[i * B * Lq:(i + 1) * B * Lq], vp["w_ip"], x0, grp=(Lq, T, r0), add_tab=pos, add_off=r0 + 2)

    # ---- forward ------------------------------------------------------------------------
    def _forward_impl(self, text_tokens: torch.Tensor, images: torch.Tensor, logits: torch.Tensor | None = None,
                      img_rows=(2,), keep: dict | None = None):
        """images: (B*m, 3, H, W) fp32 (or raw uint8, planar / channels-last) in (sequence, image) order, m = len(img_rows)."""
        cfg = self.cfg
        B, t_text = text_tokens.shape
        Lq, m = cfg.p_latents, len(img_rows)
        T = t_text + Lq * m
        dp = self.decoder._pack()
        x0 = self._ws.get("x0", (B * T, cfg.dim), torch.float32, text_tokens.device)
        xv = self._vit(images, media=m)
        self._perceive_project(xv, B, x0, T, img_rows)
        ops.embed_splice_pos(text_tokens, dp["embed"], dp["pos"], x0, img_rows=img_rows, n_img=Lq, err_flag=self._err_flag(),
                             alias_positions=cfg.alias_embed_positions)
        if keep is not None:
            keep["vit"] = xv.view(-1, cfg.vit_tokens, cfg.vit_dim).clone()
            keep["x0"] = x0.view(B, T, cfg.dim).clone()
            keep["inner_states"] = [keep["x0"]]
        if logits is None:
            logits = self._new_logits(B * T, x0.device)
        return self.decoder.run_layers(x0, B, T, logits, keep=keep)

    @torch.no_grad()
    def stages(self, text_tokens, images, image_positions=None, precision: str | None = None) -> dict:
        """Intermediate tensors of one forward (the oracle's ``stages`` keys: vit, x0, inner_states, logits) for per-stage
        parity tables; eager, never graphed."""
        text_tokens, images, img_rows, T = self._prepare_inputs(text_tokens, images, image_positions)
        keep = {}
        if (precision or self.precision) == "bf16x3":
            logits = self._accurate().forward(text_tokens, images, img_rows, keep=keep)
        else:
            logits = self._forward_impl(text_tokens, images, img_rows=img_rows, keep=keep)
        keep["logits"] = logits.view(text_tokens.shape[0], T, -1)[..., :self.cfg.vocab]
        return keep

    def _new_logits(self, M, device):
        """fp32 [M, vocab], or bf16 [M, vocab rounded up to 8] (16-byte row pitch: the TMA-store epilogue applies)."""
        if self.logits_dtype == torch.float32:
            return torch.empty(M, self.cfg.vocab, dtype=torch.float32, device=device)
        return torch.empty(M, (self.cfg.vocab + 7) // 8 * 8, dtype=torch.bfloat16, device=device)

    def _err_flag(self):
        f = getattr(self, "_errf", None)
        if f is None or not f.is_cuda:
            f = torch.zeros(1, dtype=torch.int32, device=self.embed.weight.device)
            self._errf = f
        return f

    def _prepare_inputs(self, text_tokens, images, image_positions, extra_rows=0, normalize_images=False):
        """Argument checks shared by forward and generate -> (text_tokens, images, img_rows, T).  images come back as
        (B*m,3,H,W) fp32, or, with ``normalize_images``, as the caller's uint8 pixels (B*m,3,H,W) / (B*m,H,W,3)."""
        cfg = self.cfg
        _require_cuda(text_tokens, "text_tokens")
        _require_cuda(images, "images")
        tr = getattr(self, "_trainer", None)
        if tr is not None:                # an external torch.optim stepped the fp32 masters since the copies were staged?
            tr.refresh_if_stepped_externally()
        if text_tokens.dtype != torch.int64 or text_tokens.ndim != 2:
            raise TypeError("text_tokens must be an int64 tensor of shape (B, T_text)")
        planar = (3, cfg.image, cfg.image)
        if normalize_images:
            if images.dtype != torch.uint8:
                raise TypeError("normalize_images=True takes raw uint8 pixels (CLIP rescale + normalise run on the device)")
            shapes = (planar, (cfg.image, cfg.image, 3))
            if images.ndim in (4, 5) and tuple(images.shape[-3:]) not in shapes and (images.shape[-1] == 3 or images.shape[-3] == 3):
                # raw pictures of another size: CLIPImageProcessor's resize + centre crop first (kx_resize_crop_u8)
                from .preprocess import resize_center_crop_u8
                lead = images.shape[:-3]
                images = resize_center_crop_u8(images.reshape(-1, *images.shape[-3:]), cfg.image, cfg.image)
                images = images.view(*lead, cfg.image, cfg.image, 3)
        else:
            shapes = (planar,)
        if images.ndim not in (4, 5) or tuple(images.shape[-3:]) not in shapes:
            raise ValueError(f"Input image size ({tuple(images.shape[1:])}) doesn't match model "
                             f"(3, {cfg.image}, {cfg.image}).")
        if images.shape[0] != text_tokens.shape[0]:
            raise ValueError("text_tokens and images must have the same batch size")
        m = images.shape[1] if images.ndim == 5 else 1
        pos = [2] if image_positions is None else [int(p) for p in image_positions]
        if len(pos) != m or m < 1 or m > _abi.KX_MAX_IMAGES:
            raise ValueError(f"{m} images per sequence need {m} image_positions (at most {_abi.KX_MAX_IMAGES}), got {pos}")
        if sorted(pos) != pos or pos[0] < 0 or pos[-1] > text_tokens.shape[1]:
            raise ValueError(f"image_positions {pos} must be ascending text-token indices in [0, {text_tokens.shape[1]}]")
        if image_positions is None and text_tokens.shape[1] < 2:
            raise ValueError("text_tokens needs at least 2 tokens (image features are spliced after token 1)")
        img_rows = tuple(p + i * cfg.p_latents for i, p in enumerate(pos))     # first spliced row of each image
        T = text_tokens.shape[1] + cfg.p_latents * m
        if T + extra_rows + 2 > cfg.max_positions:
            raise ValueError(f"spliced sequence length {T + extra_rows} exceeds the positional table: max is "
                             f"{cfg.max_positions - 2} (construct Kosmos(max_positions=...) to extend it)")
        if normalize_images:
            images = images.reshape(-1, *images.shape[-3:]).contiguous()
        else:   # HF casts pixels to the weight dtype ([HF]:208-209)
            images = images.to(torch.float32).reshape(-1, 3, cfg.image, cfg.image).contiguous()
        return text_tokens.contiguous(), images, img_rows, T

    def forward(self, text_tokens: torch.Tensor, images: torch.Tensor, image_positions=None, normalize_images: bool = False,
                precision: str | None = None, **kwargs):
        """Reference call (model.py:208-253): images (B,3,H,W), features spliced in front of text token 2.
        Extension (BASELINE.json configs[4]): images (B,m,3,H,W) with ``image_positions`` = m ascending text-token
        indices; image i's 64 feature rows are spliced in front of text token image_positions[i].
        ``normalize_images=True``: images are raw uint8 pixels of the model's size, planar or
        channels-last; the CLIPImageProcessor rescale + normalise that ``KosmosTokenizer.tokenize_images`` would have
        applied on the host (model.py:81-97) runs inside the patch-pack kernel.  Without it a uint8 tensor is cast
        to float as HF does ([HF]:208-209)."""
        if not isinstance(text_tokens, torch.Tensor) or not isinstance(images, torch.Tensor):
            raise TypeError("text_tokens and images must be instances of torch.Tensor")
        cfg = self.cfg
        try:
            text_tokens, images, img_rows, T = self._prepare_inputs(text_tokens, images, image_positions,
                                                                    normalize_images=normalize_images)
        except Exception as e:
            log.error(f"Failed during input validation: {e}")
            raise
        try:
            B = text_tokens.shape[0]
            if self.training and torch.is_grad_enabled():
                return self._forward_train(text_tokens, images, img_rows)
            precision = precision or self.precision
            if precision not in self._PRECISIONS:
                raise ValueError(f"precision must be one of {self._PRECISIONS}")
            if precision == "bf16x3":
                if images.dtype == torch.uint8:          # raw pixels: the bit-exact device normalise, then fp32 from there on
                    images = ops.clip_normalize_u8(images, image=cfg.image)
                return self._accurate().forward(text_tokens, images, img_rows).view(B, T, cfg.vocab)
            if self.cuda_graph:
                logits = self._forward_graphed(text_tokens, images, img_rows)
                if not self.graph_alias_output:
                    logits = logits.clone()              # the caller owns the result (reference semantics)
            else:
                logits = self._forward_impl(text_tokens, images, img_rows=img_rows)
            return logits.view(B, T, -1)[..., :cfg.vocab]
        except Exception as e:
            log.error(f"Failed during model forward pass: {e}")
            raise

    def _forward_train(self, text_tokens, images, img_rows):
        """``model.train()`` with grad enabled (the reference's training loop, train.py:640-657: forward, a loss written
        in PyTorch, ``loss.backward()``, any ``torch.optim`` optimizer): the training forward of ``KosmosTrainer`` (keeps
        the activations backward needs), returned as ONE autograd node whose backward is the hand-scheduled sm_100a
        backward pass.  Gradients land in ``param.grad`` (views of the trainer's flat buffer) and accumulate across
        backward calls like autograd's do until ``zero_grad()``; call backward once per forward; the CLIP tower and the multiway ``.B`` branches are
        frozen; dropout follows the trainer's settings (the config's 0.1 / 0.1 unless overridden).  A ``KosmosTrainer`` built on this model beforehand is used, else one is made
        (``KosmosTrainer(model)``: flat fp32 master / gradient buffers)."""
        tr = getattr(self, "_trainer", None)
        if tr is None:
            from .train import KosmosTrainer
            tr = KosmosTrainer(self)
        return tr.autograd_forward(text_tokens, images, img_rows)

    @torch.no_grad()
    def generate(self, text_tokens: torch.Tensor, images: torch.Tensor, max_new_tokens: int, image_positions=None,
                 forced_tokens=None, return_logits: bool = False, cuda_graph: bool = True, one_kernel=None,
                 normalize_images: bool = False):
        """Greedy continuation (the reference stops at logits, torchscale's ``incremental_state`` is the
        decoding path it would use): vision tower -> resampler -> splice -> prompt pass with KV-cache fill, then one-token
        steps on the weight-streaming kernels.  Returns int64 (B, max_new_tokens) on the device (and the (B, n, vocab)
        logits with ``return_logits``).  New tokens continue the spliced sequence: positions T+2, T+3, ..."""
        if not isinstance(text_tokens, torch.Tensor) or not isinstance(images, torch.Tensor):
            raise TypeError("text_tokens and images must be instances of torch.Tensor")
        text_tokens, images, img_rows, T = self._prepare_inputs(text_tokens, images, image_positions, int(max_new_tokens),
                                                                normalize_images=normalize_images)
        cfg = self.cfg
        B, m = text_tokens.shape[0], len(img_rows)
        dp = self.decoder._pack()
        x0 = self._ws.get("x0", (B * T, cfg.dim), torch.float32, text_tokens.device)
        xv = self._vit(images, media=m)
        self._perceive_project(xv, B, x0, T, img_rows)
        ops.embed_splice_pos(text_tokens, dp["embed"], dp["pos"], x0, img_rows=img_rows, n_img=cfg.p_latents,
                             err_flag=self._err_flag(), alias_positions=cfg.alias_embed_positions)
        # a generated token continues the TEXT: under alias_embed_positions it takes its text-index position too, so that
        # a step equals Kosmos.forward over the grown text (what a user of the reference, which has no generate, would run)
        text_off = cfg.p_latents * m if cfg.alias_embed_positions else -1
        return self._generate(x0, B, T, int(max_new_tokens), forced_tokens, return_logits, cuda_graph, one_kernel, text_off)

    def check_tokens(self):
        """Host-side check (one sync) that no token id of any forward so far was out of range."""
        if getattr(self, "_errf", None) is not None and int(self._errf.item()) != 0:
            self._errf.zero_()
            raise ValueError(f"token id out of range [0, {self.cfg.vocab})")

    # ---- CUDA graph replay ----------------------------------------------------------------
    def _forward_graphed(self, text_tokens, images, img_rows=(2,)):
        """One captured graph per input shape and per output slot.  Two output slots alternate, so the
        logits returned by a call stay valid until the second-next call with the same shapes (lets a
        caller overlap a device->host copy of the result with the next forward)."""
        key = (tuple(text_tokens.shape), tuple(images.shape), images.dtype, tuple(img_rows))
        g = self._graphs.get(key)
        if g is None:
            st_tok, st_img = text_tokens.clone(), images.clone()
            B, t_text = text_tokens.shape
            M = B * (t_text + self.cfg.p_latents * len(img_rows))
            outs = [self._new_logits(M, text_tokens.device) for _ in range(2)]
            self._forward_impl(st_tok, st_img, outs[0], img_rows)   # warm-up: stages weights, allocates workspaces
            torch.cuda.synchronize()
            graphs = []
            for o in outs:
                graph = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(graph):
                    self._forward_impl(st_tok, st_img, o, img_rows)
                nodes = ops.launch_count() - n0
                graphs.append(graph)
            g = [graphs, st_tok, st_img, outs, 0, nodes]
            self._graphs[key] = g
        graphs, st_tok, st_img, outs, slot, nodes = g
        g[4] = slot ^ 1
        st_tok.copy_(text_tokens)
        st_img.copy_(images)
        graphs[slot].replay()
        ops.count_graph_replay(nodes)
        return outs[slot]


class KosmosLanguage(_KosmosBase):
    """``kosmosx.model.KosmosLanguage`` (reference model.py:256-320): the text-only decoder.
    ``alibi_pos_bias`` / ``alibi_num_heads`` are accepted and ignored, as torchscale's DecoderConfig
    ignores them."""

    def __init__(self, vocab_size: int = 64007, dim: int = 2048, depth: int = 24, ffn_dim: int = 8192,
                 dropout: float = 0.1, multiway: bool = True, decoder_heads: int = 32, activation_fn: str = "gelu",
                 subln: bool = True, alibi_pos_bias: bool = True, alibi_num_heads: int = 16, xpos_rel_pos: bool = True,
                 max_rel_pos: int = 2048, *args, device=None, max_positions: int | None = None, precision: str = "bf16",
                 **kwargs):
        super().__init__()
        if precision not in self._PRECISIONS:
            raise ValueError(f"precision must be one of {self._PRECISIONS}")
        self.precision = precision
        if activation_fn != "gelu" or not subln or not xpos_rel_pos:
            raise NotImplementedError("the B200 build implements the reference configuration: gelu, subln, xpos")
        cfg = KosmosConfig(vocab=vocab_size, dim=dim, layers=depth, ffn=ffn_dim, heads=decoder_heads, multiway=multiway,
                           max_positions=max_positions or dim)          # PositionalEmbedding(dim, dim, 1), model.py:281
        cfg.validate()
        self.cfg = cfg
        with torch.device(device) if device is not None else _nullctx():
            self.embed = nn.Embedding(vocab_size, dim, padding_idx=1)
            _embed_init(self.embed)
            self.embed_positions = nn.Embedding(cfg.max_positions, dim, padding_idx=1)
            self.output_projection = nn.Linear(dim, vocab_size, bias=False)
            self.config = cfg
            self.decoder = Decoder(cfg, self.embed, self.embed_positions, self.output_projection)
        self._ws = _Workspace()
        self.eval()

    def forward(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        if not isinstance(x, torch.Tensor):
            raise TypeError("x must be a torch.Tensor of token ids")
        _require_cuda(x, "x")
        if x.dtype != torch.int64 or x.ndim != 2:
            raise TypeError("x must be an int64 tensor of shape (B, T)")
        B, T = x.shape
        if T + 2 > self.cfg.max_positions:
            raise ValueError(f"sequence length {T} exceeds the positional table: max is {self.cfg.max_positions - 2}")
        if (kwargs.get("precision") or self.precision) == "bf16x3":
            return self._accurate().forward_language(x).view(B, T, self.cfg.vocab)
        dp = self.decoder._pack()
        x0 = self._ws.get("x0", (B * T, self.cfg.dim), torch.float32, x.device)
        ops.embed_splice_pos(x.contiguous(), dp["embed"], dp["pos"], x0)
        return self.decoder.run_layers(x0, B, T).view(B, T, self.cfg.vocab)

    @torch.no_grad()
    def generate(self, x: torch.Tensor, max_new_tokens: int, forced_tokens=None, return_logits: bool = False,
                 cuda_graph: bool = True, one_kernel=None):
        """Greedy continuation of the token prefix x (B, T) through the KV-cache path (see Kosmos.generate)."""
        _require_cuda(x, "x")
        if x.dtype != torch.int64 or x.ndim != 2:
            raise TypeError("x must be an int64 tensor of shape (B, T)")
        B, T = x.shape
        if T + int(max_new_tokens) + 2 > self.cfg.max_positions:
            raise ValueError(f"sequence length {T + int(max_new_tokens)} exceeds the positional table: max is "
                             f"{self.cfg.max_positions - 2}")
        dp = self.decoder._pack()
        x0 = self._ws.get("x0", (B * T, self.cfg.dim), torch.float32, x.device)
        ops.embed_splice_pos(x.contiguous(), dp["embed"], dp["pos"], x0)
        return self._generate(x0, B, T, int(max_new_tokens), forced_tokens, return_logits, cuda_graph, one_kernel)


class _nullctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class KosmosTokenizer:
    """``kosmosx.model.KosmosTokenizer`` (reference model.py:23-129): host-side preprocessing, SURVEY.md §8(f)4.

    ``KosmosTokenizer()`` does what the reference does: it loads the CLIP processor and the gpt-neox tokenizer from the
    hub (model.py:36-46) and re-raises, after ``logging.error``, when that fails (no network here).  Keyword-only
    extras for offline use: ``tokenizer`` / ``processor`` inject already-built objects with the HF call conventions
    (``tokenizer(texts, return_tensors="pt", padding=True, truncation=True).input_ids``,
    ``tokenizer.convert_tokens_to_ids``, ``tokenizer.pad_token_id``, ``processor(images=..., return_tensors="pt")
    .pixel_values``).  Attributes as in the reference: ``processor``, ``tokenizer``, ``im_idx``, ``im_end_idx``.

    B200 addition: ``tokenize_images`` given a uint8 CUDA tensor ((N,3,H,W) or (N,H,W,3), any size) runs the processor's
    whole pipeline on the device: shortest-edge bicubic resize + centre crop (``kx_resize_crop_u8``, PIL's fixed-point
    resampling bit for bit) when the size differs from the model's, then rescale + normalise (``kx_clip_normalize_u8``),
    and returns the same fp32 ``pixel_values``; ``Kosmos.forward(..., normalize_images=True)`` fuses the last step into
    the patch pack instead.  Anything else (PIL images, lists) goes through the injected processor on the host.
    """

    CLIP_REPO = "laion/CLIP-ViT-L-14-laion2B-s32B-b82K"      # model.py:37
    TEXT_REPO = "EleutherAI/gpt-neox-20b"                    # model.py:40
    N_IMAGE_FEATURES = 64                                    # model.py:117 (dummy_image_features)

    def __init__(self, *, tokenizer=None, processor=None, image_size: int = 224):
        try:
            if processor is None:
                from transformers import CLIPProcessor
                processor = CLIPProcessor.from_pretrained(self.CLIP_REPO)
            if tokenizer is None:
                from transformers import AutoTokenizer
                tokenizer = AutoTokenizer.from_pretrained(
                    self.TEXT_REPO, additional_special_tokens=["<image>", "</image>"], eos_token="<eos>",
                    pad_token="<pad>", extra_ids=0, model_max_length=8192)
        except Exception as e:
            log.error(f"Failed to initialize KosmosTokenizer: {e}")
            raise
        self.processor = processor
        self.tokenizer = tokenizer
        self.image_size = image_size
        self.im_idx, self.im_end_idx = self.tokenizer.convert_tokens_to_ids(["<image>", "</image>"])   # model.py:51-53

    def _image_norm(self):
        """(mean, std) of the processor in use (CLIPProcessor keeps them on ``.image_processor``)."""
        ip = getattr(self.processor, "image_processor", self.processor)
        mean, std = getattr(ip, "image_mean", None), getattr(ip, "image_std", None)
        return (tuple(mean) if mean is not None else ops.CLIP_MEAN, tuple(std) if std is not None else ops.CLIP_STD)

    def tokenize_texts(self, texts):
        """-> (tokens with ``<image> </image>`` inserted after the first token, the text tokens alone): model.py:55-80."""
        try:
            texts = self.tokenizer(texts, return_tensors="pt", padding=True, truncation=True).input_ids
            image_tokens = torch.tensor([[self.im_idx, self.im_end_idx]] * texts.shape[0], dtype=texts.dtype,
                                        device=texts.device)
            return torch.cat([texts[:, 0:1], image_tokens, texts[:, 1:]], dim=1), texts
        except Exception as e:
            log.error(f"Failed to tokenize texts: {e}")
            raise

    def tokenize_images(self, images):
        """-> fp32 ``pixel_values`` (N,3,H,W): model.py:82-97.  uint8 CUDA tensors of the model's size are normalised on
        the device; everything else is the processor's job."""
        try:
            if isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8:
                mean, std = self._image_norm()
                hw = tuple(images.shape[1:3]) if images.shape[-1] == 3 and images.shape[1] != 3 else tuple(images.shape[2:4])
                if hw != (self.image_size, self.image_size):      # any other size: the processor's resize + centre crop, on the device
                    from .preprocess import resize_center_crop_u8
                    images = resize_center_crop_u8(images, self.image_size, self.image_size)
                return ops.clip_normalize_u8(images.contiguous(), image=self.image_size, mean=mean, std=std)
            return self.processor(images=images, return_tensors="pt").pixel_values
        except Exception as e:
            log.error(f"Failed to tokenize images: {e}")
            raise

    def tokenize(self, sample):
        """{"target_text", "image"} -> {"text_tokens", "images", "labels", "attention_mask"}: model.py:99-129, including the
        reference's mask layout (64 ones in FRONT of the text mask although the features are spliced after token 1,
        SURVEY.md Appendix C item 4; ``Kosmos.forward`` ignores the mask either way)."""
        try:
            text_tokens, only_text_tokens = self.tokenize_texts(sample["target_text"])
            attention_mask = text_tokens != self.tokenizer.pad_token_id
            dummy_image_features = torch.ones((text_tokens.shape[0], self.N_IMAGE_FEATURES), device=text_tokens.device)
            attention_mask = torch.cat([dummy_image_features, attention_mask], dim=1)
            return {
                "text_tokens": text_tokens,
                "images": self.tokenize_images(sample["image"]),
                "labels": only_text_tokens,
                "attention_mask": attention_mask,
            }
        except Exception as e:
            log.error(f"Failed to tokenize sample: {e}")
            raise
