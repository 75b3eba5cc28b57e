"""Verification-precision forward ("bf16x3") of the Kosmos-X path.

BASELINE.json states the parity tolerance as logits max-abs-diff <= 1e-3 against the reference's PyTorch path
(/root/reference/kosmosx/model.py:208-253).  The throughput path rounds tensor-core operands to bf16 (one ulp at 1.0 is
7.8e-3) and lands ~4e-2 from the fp32 reference, so ``Kosmos(precision="bf16x3")`` runs the SAME tcgen05 GEMM kernel on
split operands — x = hi + lo with hi = bf16(x), lo = bf16(x - hi); one launch over K' = 3K accumulates Ah.Wh + Ah.Wl +
Al.Wh in one fp32 TMEM tile (csrc/accurate.cu) — and keeps every tensor between the GEMMs in fp32: unfused fp32
LayerNorms (kx_layernorm_fwd), fp32 GEMM epilogues (bias, erf-GELU, residual, positional add, row scatter), fp32
attention (kx_attn_f32), fp32 xPos rotation and patch im2col.  Same launch structure, same C ABI, no PyTorch arithmetic.
It exists to show the kernels compute the reference's function to the stated tolerance; bf16 stays the speed mode.
"""
from __future__ import annotations

import torch

from . import _abi, ops


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def _w3(weight: torch.Tensor, n_pad=None) -> torch.Tensor:
    """fp32 Linear weight [N, K] -> bf16 [N, 3K] in the (hi | lo | hi) weight layout."""
    w = _f32(weight)
    return ops.split_bf16x3(w.reshape(w.shape[0], -1), weights=True, n_pad=n_pad)


class _Scratch:
    def __init__(self):
        self.bufs = {}

    def get(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        b = self.bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=dtype, device=device)
            self.bufs[key] = b
        return b


class AccurateEngine:
    """Staged split weights + workspaces of one model.  ``forward`` mirrors Kosmos._forward_impl stage by stage."""

    def __init__(self, model):
        self.m = model
        self.ws = _Scratch()
        self.packed = None

    def invalidate(self):
        self.packed = None
        self.ws.bufs.clear()

    # ---- staging -----------------------------------------------------------------------------------------
    def _pack(self):
        if self.packed is not None:
            return self.packed
        from .model import _live
        m, cfg = self.m, self.m.cfg
        p = {}
        if hasattr(m, "clip_model"):
            cm = m.clip_model
            k = 3 * cfg.patch * cfg.patch
            k_pad = (k + 63) // 64 * 64
            vl = []
            for L in cm.encoder.layers:
                a = L.self_attn
                vl.append(dict(
                    ln1=(_f32(L.layer_norm1.weight), _f32(L.layer_norm1.bias)),
                    wqkv=_w3(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                    bqkv=_f32(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                    wo=_w3(a.out_proj.weight), bo=_f32(a.out_proj.bias),
                    ln2=(_f32(L.layer_norm2.weight), _f32(L.layer_norm2.bias)),
                    w1=_w3(L.mlp.fc1.weight), b1=_f32(L.mlp.fc1.bias), w2=_w3(L.mlp.fc2.weight), b2=_f32(L.mlp.fc2.bias)))
            pl = []
            for attn, ff in m.perceive.layers:
                pl.append(dict(
                    nm=(_f32(attn.norm_media.weight), _f32(attn.norm_media.bias)),
                    nl=(_f32(attn.norm_latents.weight), _f32(attn.norm_latents.bias)),
                    wq=_w3(attn.to_q.weight), wkv=_w3(attn.to_kv.weight), wout=_w3(attn.to_out.weight),
                    ff_ln=(_f32(ff[0].weight), _f32(ff[0].bias)), w1=_w3(ff[1].weight), w2=_w3(ff[3].weight)))
            p.update(k_pad=k_pad, w_patch=_w3(cm.embeddings.patch_embedding.weight, n_pad=k_pad),
                     cls=_f32(cm.embeddings.class_embedding), vpos=_f32(cm.embeddings.position_embedding.weight),
                     pre_ln=(_f32(cm.pre_layrnorm.weight), _f32(cm.pre_layrnorm.bias)), vit=vl, perceiver=pl,
                     latents=_f32(m.perceive.latents), media_pos=_f32(m.perceive.media_pos_emb).view(-1, cfg.vit_dim),
                     p_norm=(_f32(m.perceive.norm.weight), _f32(m.perceive.norm.bias)), w_ip=_w3(m.image_proj.weight))
        dec = m.decoder
        dl = []
        for L in dec.layers:
            sa = L.self_attn
            q, k, v, o = (_live(x) for x in (sa.q_proj, sa.k_proj, sa.v_proj, sa.out_proj))
            ffn = _live(L.ffn)
            ln_a, ln_i, ln_f = _live(L.self_attn_layer_norm), _live(sa.inner_attn_ln), _live(L.final_layer_norm)
            dl.append(dict(
                ln_a=(_f32(ln_a.weight), _f32(ln_a.bias)), wqkv=_w3(torch.cat([q.weight, k.weight, v.weight], 0)),
                bqkv=_f32(torch.cat([q.bias, k.bias, v.bias], 0)),
                ln_i=(_f32(ln_i.weight), _f32(ln_i.bias)), wo=_w3(o.weight), bo=_f32(o.bias),
                ln_f=(_f32(ln_f.weight), _f32(ln_f.bias)), w1=_w3(ffn.fc1.weight), b1=_f32(ffn.fc1.bias),
                ln_ffn=(_f32(ffn.ffn_layernorm.weight), _f32(ffn.ffn_layernorm.bias)), w2=_w3(ffn.fc2.weight), b2=_f32(ffn.fc2.bias)))
        p.update(dec=dl, ln_out=(_f32(dec.layer_norm.weight), _f32(dec.layer_norm.bias)),
                 w_out=_w3(dec.output_projection.weight),
                 b_out=None if dec.output_projection.bias is None else _f32(dec.output_projection.bias),
                 embed=_f32(dec.embed_tokens.weight), pos=_f32(dec.embed_positions.weight))
        self.packed = p
        return p

    # ---- building blocks ---------------------------------------------------------------------------------
    def _gemm(self, a: torch.Tensor, w3: torch.Tensor, out: torch.Tensor, n_pad=None, **kw):
        """out = epilogue(a . W^T) at split precision: a fp32 [M, K] -> (hi | hi | lo) scratch, one tcgen05 launch over 3K."""
        n_pad = a.shape[1] if n_pad is None else n_pad
        a3 = self.ws.get(f"a3_{a.shape[0]}x{n_pad}", (a.shape[0], 3 * n_pad), torch.bfloat16, a.device)
        ops.split_bf16x3(a, a3, weights=False, n_pad=n_pad)
        return ops.gemm(a3, w3, out, **kw)

    def _ln(self, x, gb, out, **kw):
        return ops.layernorm(x, gb[0], gb[1], out, eps=self.m.cfg.eps, **kw)

    # ---- stages --------------------------------------------------------------------------------------------
    def vit(self, images: torch.Tensor, media: int = 1, keep=None) -> torch.Tensor:
        cfg, p, ws = self.m.cfg, self._pack(), self.ws
        f32 = torch.float32
        N = images.shape[0]
        Tv, Dv, P = cfg.vit_tokens, cfg.vit_dim, cfg.vit_tokens - 1
        M = N * Tv
        dev = images.device
        patches = ws.get("patches", (N * P, p["k_pad"]), f32, dev)
        emb = ws.get("vemb", (M, Dv), f32, dev)
        x = ws.get("vx", (M, Dv), f32, dev)
        h = ws.get("vh", (M, Dv), f32, dev)
        qkv = ws.get("vqkv", (M, 3 * Dv), f32, dev)
        att = ws.get("vatt", (M, Dv), f32, dev)
        mid = ws.get("vmid", (M, cfg.vit_mlp), f32, dev)
        ops.im2col_patches_f32(images, patches, p["cls"], p["vpos"], emb.view(N, Tv, Dv), image=cfg.image, patch=cfg.patch, media=media)
        self._gemm(patches, p["w_patch"], emb, n_pad=p["k_pad"], grp=(P, Tv, 1), add_tab=p["vpos"], add_off=1)
        self._ln(emb, p["pre_ln"], x)
        act = _abi.KX_ACT_GELU if cfg.vit_act == "gelu" else _abi.KX_ACT_QUICK_GELU
        scale = (Dv // cfg.vit_heads) ** -0.5
        for L in p["vit"]:
            self._ln(x, L["ln1"], h)
            self._gemm(h, L["wqkv"], qkv, bias=L["bqkv"])
            ops.attention_f32(qkv[:, :Dv], qkv[:, Dv:2 * Dv], qkv[:, 2 * Dv:], att, batch=N, heads=cfg.vit_heads, n_q=Tv, n_kv=Tv,
                              causal=False, scale=scale)
            self._gemm(att, L["wo"], x, bias=L["bo"], res=x)
            self._ln(x, L["ln2"], h)
            self._gemm(h, L["w1"], mid, bias=L["b1"], act=act)
            self._gemm(mid, L["w2"], x, bias=L["b2"], res=x)
        if keep is not None:
            keep["vit"] = x.view(N, Tv, Dv).clone()
        return x

    def perceive_project(self, xv, B, x0, T, img_rows, keep=None):
        cfg, p, ws = self.m.cfg, self._pack(), self.ws
        f32 = torch.float32
        Tv, Dv, Lq, Hp = cfg.vit_tokens, cfg.vit_dim, cfg.p_latents, cfg.p_heads
        inner = Hp * cfg.p_dim_head
        dev = xv.device
        m = len(img_rows)
        N = B * m
        lat = ws.get("plat", (N * Lq, Dv), f32, dev)
        cat = ws.get("pcat", (N * (Tv + Lq), Dv), f32, dev)
        lnl = ws.get("plnl", (N * Lq, Dv), f32, dev)
        q = ws.get("pq", (N * Lq, inner), f32, dev)
        kv = ws.get("pkv", (N * (Tv + Lq), 2 * inner), f32, dev)
        att = ws.get("patt", (N * Lq, inner), f32, dev)
        mid = ws.get("pmid", (N * Lq, Dv * cfg.p_ff_mult), f32, dev)
        ops.broadcast_rows(p["latents"], lat, N)
        mp = p["media_pos"][0:m]
        for L in p["perceiver"]:
            self._ln(xv, L["nm"], cat, pre_add=mp, pre_add_group=Tv * B if m > 1 else 0, grp=(Tv, Tv + Lq, 0))
            self._ln(lat, L["nl"], cat, grp=(Lq, Tv + Lq, Tv))
            self._ln(lat, L["nl"], lnl)
            self._gemm(lnl, L["wq"], q)
            self._gemm(cat, L["wkv"], kv)
            ops.attention_f32(q, kv[:, :inner], kv[:, inner:], att, batch=N, heads=Hp, n_q=Lq, n_kv=Tv + Lq, causal=False,
                              scale=cfg.p_dim_head ** -0.5)
            self._gemm(att, L["wout"], lat, res=lat)
            self._ln(lat, L["ff_ln"], lnl)
            self._gemm(lnl, L["w1"], mid, act=_abi.KX_ACT_GELU)
            self._gemm(mid, L["w2"], lat, res=lat)
        self._ln(lat, p["p_norm"], lnl)
        if keep is not None:
            keep["perceive"] = lnl.view(m, B, Lq, Dv).transpose(0, 1).clone()       # (B, m, 64, Dv)
        for i, r0 in enumerate(img_rows):
            self._gemm(lnl[i * B * Lq:(i + 1) * B * Lq], p["w_ip"], x0, grp=(Lq, T, r0), add_tab=p["pos"], add_off=r0 + 2)

    def decoder_layers(self, x: torch.Tensor, B: int, T: int, keep=None) -> torch.Tensor:
        """x fp32 [B*T, D], updated in place -> fp32 logits [B*T, vocab]."""
        cfg, p, ws = self.m.cfg, self._pack(), self.ws
        f32 = torch.float32
        M, D, F, H = B * T, cfg.dim, cfg.ffn, cfg.heads
        dev = x.device
        h = ws.get("h", (M, D), f32, dev)
        qkv = ws.get("qkv", (M, 3 * D), f32, dev)
        att = ws.get("att", (M, D), f32, dev)
        u = ws.get("u", (M, F), f32, dev)
        g = ws.get("g", (M, F), f32, dev)
        tabs = self.m.decoder._xpos(T, dev)
        scale = (D // H) ** -0.5
        states = [x.view(B, T, D).clone()] if keep is not None else None
        for L in p["dec"]:
            self._ln(x, L["ln_a"], h)
            self._gemm(h, L["wqkv"], qkv, bias=L["bqkv"])
            ops.xpos_apply_f32(qkv, D, T, tabs)
            ops.attention_f32(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, batch=B, heads=H, n_q=T, n_kv=T, causal=True, scale=scale)
            self._ln(att, L["ln_i"], h)
            self._gemm(h, L["wo"], x, bias=L["bo"], res=x)
            self._ln(x, L["ln_f"], h)
            self._gemm(h, L["w1"], u, bias=L["b1"], act=_abi.KX_ACT_GELU)
            self._ln(u, L["ln_ffn"], g)
            self._gemm(g, L["w2"], x, bias=L["b2"], res=x)
            if states is not None:
                states.append(x.view(B, T, D).clone())
        if keep is not None:
            keep["inner_states"] = states
        self._ln(x, p["ln_out"], h)
        logits = torch.empty(M, p["w_out"].shape[0], dtype=f32, device=dev)
        self._gemm(h, p["w_out"], logits, bias=p["b_out"])
        return logits

    def forward(self, text_tokens, images, img_rows, keep=None):
        m, cfg, p = self.m, self.m.cfg, self._pack()
        B, t_text = text_tokens.shape
        Lq, nm = cfg.p_latents, len(img_rows)
        T = t_text + Lq * nm
        x0 = self.ws.get("x0", (B * T, cfg.dim), torch.float32, text_tokens.device)
        xv = self.vit(images, media=nm, keep=keep)
        self.perceive_project(xv, B, x0, T, img_rows, keep=keep)
        ops.embed_splice_pos(text_tokens, p["embed"], p["pos"], x0, img_rows=img_rows, n_img=Lq, err_flag=m._err_flag(),
                             alias_positions=cfg.alias_embed_positions)
        if keep is not None:
            keep["x0"] = x0.view(B, T, cfg.dim).clone()
        return self.decoder_layers(x0, B, T, keep=keep)

    def forward_language(self, tokens):
        cfg, p = self.m.cfg, self._pack()
        B, T = tokens.shape
        x0 = self.ws.get("x0", (B * T, cfg.dim), torch.float32, tokens.device)
        ops.embed_splice_pos(tokens.contiguous(), p["embed"], p["pos"], x0)
        return self.decoder_layers(x0, B, T)
