"""Data-parallel training step of the Kosmos-X path on the sm_100a kernels.

Shape of the reference's step (train.py:643-657): forward -> loss -> backward -> clip_grad_norm_(1.0) ->
optimizer.step() -> zero_grad(), with AdamW or Lion over a decay / no-decay split (train.py:257-398) and one
gradient all-reduce per step across the data-parallel ranks.  The reference script itself cannot run
(`model(inputs, return_loss=True)` at train.py:647 passes no images and Kosmos returns no loss — SURVEY.md
fact 8); the loss is the intended one of experimental/model/allModalities/notes.txt:566-574: next-token
cross-entropy over the text rows of the spliced sequence.

Everything arithmetic is a C-ABI call (ops.*): the training forward keeps what backward needs (LayerNorm outputs
are materialised instead of folded, the FFN keeps its pre-activation), backward is hand-scheduled — no autograd.
Gradients live in one flat fp32 buffer that is all-reduced in per-layer buckets while backward is still running.

Trained here (SURVEY.md §8(e) trainable set): the decoder (.A branches), the final LayerNorm, the LM head, the
token-embedding and position tables, the perceiver resampler and image_proj.  Frozen: the CLIP tower (the reference's
notes.txt:537 `clip_model.requires_grad_(False)`; its optional last-layer fine-tuning is not built) and the multiway
.B branches (never executed).  Dropout (dropout = attention_dropout = 0.1 in the reference's train mode,
model.py:175-177) is applied at torchscale's four sites: the decoder input, out_proj's output, fc2's output (Philox masks
regenerated in backward) and the attention probabilities (keep bits drawn ahead of the two flash kernels, 1 bit per score).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _abi, ops
from .model import Kosmos, _live

_ALIGN = 64          # elements: keeps every parameter 256-byte aligned in the fp32 buffers (128 in the bf16 copy)


def _round_up(n, a=_ALIGN):
    return (n + a - 1) // a * a


class _Seg:
    __slots__ = ("off", "shape", "numel")

    def __init__(self, off, shape):
        self.off, self.shape = off, tuple(shape)
        self.numel = 1
        for s in self.shape:
            self.numel *= s


class KosmosTrainer:
    """One object per process (one process per GPU).  ``step(text_tokens, images)`` runs a whole optimisation step
    and returns the mean loss as a device scalar (no host sync)."""

    SITE_X0 = 0xFFFF0000          # dropout site ids: layer * 4 + {0: out_proj output, 1: fc2 output, 2: attention probabilities}

    def __init__(self, model: Kosmos, *, optimizer: str = "adamw", lr: float = 1e-4, betas=(0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.1, max_grad_norm: float = 1.0, process_group=None, overlap_all_reduce: bool = False,
                 train_resampler: bool = True, layout_only: bool = False, loss_rule: str = "reference",
                 pad_token_id: int | None = None, lr_schedule=None, grad_reduce_dtype: torch.dtype = torch.bfloat16,
                 distributed: bool | None = None, dropout: float | None = None, attention_dropout: float | None = None,
                 seed: int = 0, bwd_max_ctas: int = 0, recompute: bool = False, train_clip_last_layer: bool = False,
                 shard_optimizer: bool = False):
        """overlap_all_reduce: False (default) = ONE all-reduce of the whole flat gradient buffer after backward; True = per-layer
        buckets issued while backward is still running.  Measured on 2 and 8 B200s (profiles/r2_nccl_overlap.md): NCCL's kernels
        occupy SMs that the persistent, one-CTA-per-SM backward kernels are sized for, so every GEMM / LayerNorm-backward
        launch that overlaps a bucket runs a partial second wave — the overlapped exchange costs MORE (+8 ms at 2 GPUs, +16 ms
        at 8) than its exposed time (2.7 GB of bf16 over NVLink: ~6 ms).
        loss_rule: "reference" = the rows / targets of the reference's intended loop (notes.txt:566-574: the `<image>`
        `</image>` markers and the feature rows carry no loss and are never targets; row 0 predicts the first real text
        token), "next_token" = plain shift by one over the text rows.  pad_token_id: targets equal to it are ignored
        (None = the reference's loop, which masks nothing).  lr_schedule: callable step -> multiplier of ``lr`` (see
        ``cosine_with_warmup``), evaluated on the host from the step COUNT — no device value is read.
        grad_reduce_dtype: torch.bfloat16 (default) exchanges a bf16 copy of the gradients (2.7 GB instead of 5.4 GB; the
        reference's FSDP ``reduce_dtype`` is 16-bit too, train.py:156-162) and converts the received sums back to fp32;
        torch.float32 all-reduces the flat fp32 buffer itself (bit-exact sum of the ranks' fp32 gradients).  distributed: None = data parallel over ``process_group`` (or the
        default group) whenever torch.distributed is initialised; False = this process trains alone.
        dropout / attention_dropout: None = the model's config (the reference trains with 0.1 / 0.1, model.py:175-177);
        0 switches a site off (parity runs).  Masks are Philox4x32-7 functions of (seed, forward count, site, coordinates):
        the element-wise sites (decoder input, out_proj output, fc2 output) are regenerated in backward, the attention
        probabilities' keep bits (1 bit per score) are drawn by kx_attn_dropout_masks ahead of the flash kernels.
        Ranks draw different masks (the seed is offset by the rank).
        recompute: activation checkpointing per decoder layer (the reference wraps the decoder with torch's checkpoint_wrapper,
        train.py:84-110,528-529): the forward keeps only each layer's fp32 input (134 MB per layer at C3 instead of 1.28 GB) and
        backward re-runs the layer before differentiating it — bit-identical gradients for one more decoder forward per step.
        train_clip_last_layer: also train the last encoder layer of the CLIP ViT (the reference freezes CLIP except its last
        layer, notes.txt:537-538 / model.py:184-190); the other 23 stay frozen on the inference kernels.  Needs train_resampler
        (the gradient reaches the ViT through the resampler's norm_media).
        shard_optimizer: with more than one rank, ``step`` / ``step_accumulated`` run the optimizer ZeRO-1 style: the bf16 gradients
        of the weight-decay segment (every Linear weight, 1.3 G of the 1.37 G trained parameters) are REDUCE-SCATTERED instead of
        all-reduced, each rank clips (its shard's sum of squares + one scalar all-reduce) and updates only its 1/world slice of the
        fp32 masters / moments, and the updated bf16 tensor-core copies are ALL-GATHERED — the same bytes on NVLink as the one
        all-reduce (which is a reduce-scatter + all-gather inside NCCL), the 8 ms optimizer pass divided by the world size, no
        cast of the whole buffer back to fp32.  The small no-decay segment (biases, LayerNorms, the embedding tables) stays
        replicated.  The fp32 masters and the moments outside a rank's slice go stale: ``gather_masters()`` (a collective, call it
        on EVERY rank) brings the masters up to date before ``state_dict()`` / ``save_checkpoint`` / an inference forward, which
        raise until then.  ``loss_and_grads`` and the autograd bridge keep the all-reduce (their contract is the summed gradient in
        ``param.grad``).  Needs the bf16 exchange and ``overlap_all_reduce=False``.
        bwd_max_ctas: > 0 caps the grid of the persistent backward GEMMs (they are sized to one CTA per SM; an NCCL kernel
        resident on a few SMs while they launch would push the CTAs that do not fit into a second wave)."""
        if optimizer not in ("adamw", "lion"):
            raise ValueError("optimizer must be 'adamw' or 'lion' (train.py:375-386)")
        if loss_rule not in ("reference", "next_token"):
            raise ValueError("loss_rule must be 'reference' or 'next_token'")
        if grad_reduce_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("grad_reduce_dtype must be torch.float32 or torch.bfloat16")
        self.model = model
        self.cfg = model.cfg
        self.opt, self.lr, self.betas, self.eps, self.wd = optimizer, lr, betas, eps, weight_decay
        self.loss_rule, self.pad_token_id, self.lr_schedule = loss_rule, pad_token_id, lr_schedule
        self.p_drop = float(model.cfg.dropout if dropout is None else dropout)
        self.p_attn = float(model.cfg.attention_dropout if attention_dropout is None else attention_dropout)
        if not (0.0 <= self.p_drop < 1.0 and 0.0 <= self.p_attn < 1.0):
            raise ValueError("dropout probabilities must be in [0, 1)")
        self.seed = int(seed)
        self._fw_count = 0
        self.bwd_max_ctas = int(bwd_max_ctas)
        self.recompute = bool(recompute)
        self.grad_reduce_dtype = grad_reduce_dtype
        self.max_grad_norm = max_grad_norm
        self.pg = process_group
        self.overlap = overlap_all_reduce
        self.train_resampler = train_resampler
        self.train_clip_last = bool(train_clip_last_layer)
        if self.train_clip_last and not train_resampler:
            raise ValueError("train_clip_last_layer needs train_resampler=True (the gradient reaches the ViT through the resampler)")
        self.world = 1
        if distributed is not False and (process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        self.shard_optimizer = bool(shard_optimizer) and self.world > 1
        if self.shard_optimizer and (overlap_all_reduce or grad_reduce_dtype != torch.bfloat16):
            raise ValueError("shard_optimizer needs grad_reduce_dtype=torch.bfloat16 and overlap_all_reduce=False")
        self._defer_reduce = False               # step() of the sharded optimizer: backward leaves the exchange to _optimize_sharded
        self._masters_sharded = False            # the fp32 masters outside this rank's slice are stale (see gather_masters)
        self.t = 0
        self._ws = {}
        self._fw_serial = 0                      # autograd bridge: which forward the saved activations belong to
        self._flatten(layout_only)
        if not layout_only:
            model._trainer = self                # Kosmos.forward in train mode builds its autograd graph on this trainer

    # ------------------------------------------------------------------ flat parameter / gradient buffers
    def _flatten(self, layout_only=False):
        """Lay every trained parameter out in one flat buffer: [decay segment | no-decay segment], per-layer blocks
        first.  layout_only (tests of the bucket plan on a CPU box) stops before anything touches the device."""
        m, cfg = self.model, self.cfg
        dev = m.embed.weight.device
        if dev.type != "cuda" and not layout_only:
            raise RuntimeError("KosmosTrainer: move the model to a B200 first (there is no CPU path)")
        dec = m.decoder
        decay, nodecay = [], []                 # (param, key) in buffer order
        self.layers = []
        for L in dec.layers:
            sa = L.self_attn
            q, k, v, o = (_live(x) for x in (sa.q_proj, sa.k_proj, sa.v_proj, sa.out_proj))
            ffn = _live(L.ffn)
            ln_a, ln_i, ln_f = _live(L.self_attn_layer_norm), _live(sa.inner_attn_ln), _live(L.final_layer_norm)
            decay += [q.weight, k.weight, v.weight, o.weight, ffn.fc1.weight, ffn.fc2.weight]      # q|k|v adjacent: one [3D, D] view
            nodecay += [q.bias, k.bias, v.bias, o.bias, ffn.fc1.bias, ffn.fc2.bias, ln_a.weight, ln_a.bias, ln_i.weight,
                        ln_i.bias, ln_f.weight, ln_f.bias, ffn.ffn_layernorm.weight, ffn.ffn_layernorm.bias]
            self.layers.append(dict(q=q, k=k, v=v, o=o, fc1=ffn.fc1, fc2=ffn.fc2, ln_a=ln_a, ln_i=ln_i, ln_f=ln_f,
                                    ln_ffn=ffn.ffn_layernorm))
        decay.append(m.output_projection.weight)
        nodecay += [dec.layer_norm.weight, dec.layer_norm.bias, m.embed.weight, m.embed_positions.weight]
        if m.output_projection.bias is not None:
            nodecay.append(m.output_projection.bias)
        self.n_vision_decay_params = 0
        if self.train_resampler:                 # perceiver resampler + image_proj: behind everything the decoder owns
            pv = m.perceive
            vd = []
            for attn, ff in pv.layers:
                vd += [attn.to_q.weight, attn.to_kv.weight, attn.to_out.weight, ff[1].weight, ff[3].weight]
                nodecay += [attn.norm_media.weight, attn.norm_media.bias, attn.norm_latents.weight, attn.norm_latents.bias,
                            ff[0].weight, ff[0].bias]
            vd.append(m.image_proj.weight)
            nodecay += [pv.norm.weight, pv.norm.bias, pv.latents, pv.media_pos_emb]
            self.clip_last = None
            if self.train_clip_last:             # last CLIP encoder layer: q|k|v adjacent like a decoder layer's
                Lc = m.clip_model.encoder.layers[-1]
                a = Lc.self_attn
                self.clip_last = dict(q=a.q_proj, k=a.k_proj, v=a.v_proj, o=a.out_proj, fc1=Lc.mlp.fc1, fc2=Lc.mlp.fc2,
                                      ln1=Lc.layer_norm1, ln2=Lc.layer_norm2)
                vd += [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.out_proj.weight, Lc.mlp.fc1.weight, Lc.mlp.fc2.weight]
                nodecay += [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias, a.out_proj.bias, Lc.mlp.fc1.bias, Lc.mlp.fc2.bias,
                            Lc.layer_norm1.weight, Lc.layer_norm1.bias, Lc.layer_norm2.weight, Lc.layer_norm2.bias]
                if cfg.vit_dim % _ALIGN:
                    raise ValueError("train_clip_last_layer: vit_dim must be a multiple of 64")
            decay += vd
            self.n_vision_decay_params = len(vd)
        d = cfg.dim
        for q, k, v in ((L["q"], L["k"], L["v"]) for L in self.layers):
            if q.weight.numel() % _ALIGN or q.bias.numel() % _ALIGN:
                raise ValueError("dim must be a multiple of 64")
        self.seg = {}
        off = 0
        for p in decay:
            self.seg[id(p)] = _Seg(off, p.shape)
            off += _round_up(p.numel())
        self.n_decay = off
        # where the resampler's Linear weights start inside the decay segment (they complete last, after the decoder)
        self.vision_decay_off = self.seg[id(decay[-self.n_vision_decay_params])].off if self.n_vision_decay_params else off
        for p in nodecay:
            self.seg[id(p)] = _Seg(off, p.shape)
            off += _round_up(p.numel())
        self.n_total = off
        self.params = decay + nodecay
        # sharded optimizer: the decay segment is cut into `world` equal chunks (padded; the in-place all-gather of the bf16 copies
        # and the reduce-scatter of the bf16 gradients need equal counts per rank)
        self._nd_pad = _round_up(self.n_decay, self.world * 1024) if self.shard_optimizer else self.n_decay
        if layout_only:
            return
        f32 = dict(dtype=torch.float32, device=dev)
        self.P = torch.zeros(off, **f32)
        self.G = torch.zeros(off, **f32)
        self.M1 = torch.zeros(off, **f32)
        self.M2 = torch.zeros(off, **f32) if self.opt == "adamw" else None
        self._W16p = torch.zeros(self._nd_pad, dtype=torch.bfloat16, device=dev)
        self.W16 = self._W16p[:self.n_decay]
        with torch.no_grad():
            for p in self.params:
                s = self.seg[id(p)]
                view = self.P[s.off:s.off + s.numel].view(s.shape)
                view.copy_(p.data)
                p.data = view                                            # the module now lives in the flat buffer
                p.grad = self.G[s.off:s.off + s.numel].view(s.shape)
        # every parameter of the model that is NOT trained here stays frozen
        trained = {id(p) for p in self.params}
        for p in m.parameters():
            if id(p) not in trained:
                p.requires_grad_(False)
        self.scalars = torch.zeros(8, **f32)       # [0:2] loss sum / rows, [2] sum g^2, [3] clip scale, [4] grad norm, [5] loss rows, [6:8] partial sums g^2 (sharded optimizer)
        if self.shard_optimizer:
            m.register_state_dict_pre_hook(lambda module, prefix, keep_vars: self._require_whole_masters("state_dict()"))
        self.sync_weights()

    def sync_weights(self):
        """Re-derive the bf16 tensor-core copies from the fp32 master weights (after load_state_dict or any
        in-place edit of the parameters).  The optimizer kernels keep them in sync afterwards."""
        self._require_whole_masters("sync_weights()")
        ops.cast_bf16(self.P[:self.n_decay], self.W16)
        self._p_version = self._params_version()
        self._inference_copies_stale()

    def _params_version(self):
        """Sum of PyTorch's in-place version counters of the trained parameters (each ``Parameter`` keeps its own, even
        though its storage is a slice of the flat buffer): changes whenever anything edits a parameter through PyTorch."""
        return sum(p._version for p in self.params)

    def refresh_if_stepped_externally(self):
        """An external optimizer (the autograd-bridge loop) or ``load_state_dict`` edits the fp32 masters in place, which no
        kernel of this library sees happen: PyTorch's version counters of the parameters tell.  Called by the inference entry
        points of the model, so an eval forward / generate right after ``optimizer.step()`` runs on the stepped weights."""
        self._require_whole_masters("an inference forward")
        if self._params_version() != self._p_version:
            self.sync_weights()

    # ------------------------------------------------------------------ sharded optimizer (ZeRO-1)
    def shard_range(self, rank=None):
        """[lo, hi) of the decay segment whose masters / moments ``rank`` (default: this one) owns under shard_optimizer."""
        if rank is None:
            rank = torch.distributed.get_rank(self.pg)
        chunk = self._nd_pad // self.world
        return min(rank * chunk, self.n_decay), min((rank + 1) * chunk, self.n_decay)

    def _require_whole_masters(self, what):
        if self._masters_sharded:
            raise RuntimeError(f"KosmosTrainer(shard_optimizer=True): {what} needs the whole fp32 master weights, but every rank has only "
                               "updated its own slice since the last gather — call trainer.gather_masters() on EVERY rank first")

    def gather_masters(self):
        """Collective (every rank of the group must call it): each rank broadcasts the slice of the fp32 masters it owns, so
        ``state_dict()``, ``save_checkpoint`` and the inference path see the stepped weights everywhere.  The moments stay sharded."""
        if not self._masters_sharded:
            return
        dist = torch.distributed
        for r in range(self.world):
            lo, hi = self.shard_range(r)
            if hi > lo:
                dist.broadcast(self.P[lo:hi], src=dist.get_global_rank(self.pg, r) if self.pg is not None else r, group=self.pg)
        self._masters_sharded = False

    def _inference_copies_stale(self):
        """The inference path re-stages its folded / split weights lazily from the fp32 masters."""
        m = self.model
        m.decoder._packed = None
        m._graphs = {}                           # captured forwards read the staged buffers that are about to be replaced
        if self.train_resampler:
            m._resampler_dirty = True
        if self.train_clip_last:
            m._clip_last_dirty = True
        if getattr(m, "_acc", None) is not None:
            m._acc.invalidate()

    def _w16(self, p):
        s = self.seg[id(p)]
        return self.W16[s.off:s.off + s.numel].view(s.shape)

    def _g(self, p):
        s = self.seg[id(p)]
        return self.G[s.off:s.off + s.numel].view(s.shape)

    def _qkv(self, L, what):
        """q|k|v as one [3D, D] matrix / [3D] vector (adjacent segments, no padding between them)."""
        s = self.seg[id(getattr(L["q"], what))]
        n = s.numel * 3
        if what == "weight":
            return (self.W16[s.off:s.off + n].view(3 * s.shape[0], s.shape[1]), self.G[s.off:s.off + n].view(3 * s.shape[0], s.shape[1]))
        return self.P[s.off:s.off + n], self.G[s.off:s.off + n]

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        b = self._ws.get(key)
        if b is None:
            b = torch.empty(shape, dtype=dtype, device=self.P.device)
            self._ws[key] = b
        return b

    # ------------------------------------------------------------------ forward (keeps what backward needs)
    def _forward(self, text_tokens, images, img_rows):
        m, cfg = self.model, self.cfg
        B, t_text = text_tokens.shape
        Lq = cfg.p_latents
        T = t_text + Lq * len(img_rows)
        M, D, F, H, V = B * T, cfg.dim, cfg.ffn, cfg.heads, cfg.vocab
        bf, f32 = torch.bfloat16, torch.float32
        dp = m.decoder
        # frozen vision side on the inference kernels: image rows of x0 (+ their positions)
        x = self._buf("x0", (M, D), f32)
        clip = None
        if self.train_clip_last:                 # 23 frozen layers on the inference kernels, the last one kept for backward
            clip = self._clip_last_forward(m._vit(images, media=len(img_rows), upto=-1), B * len(img_rows))
            xv = clip["xv"]
        else:
            xv = m._vit(images, media=len(img_rows))
        pos = m.embed_positions.weight
        vis = None
        if self.train_resampler:
            vis = self._resampler_forward(xv, B, x, T, img_rows, pos)
            vis["clip"] = clip
        else:
            m._perceive_project(xv, B, x, T, img_rows, pos_table=pos)
        ops.embed_splice_pos(text_tokens, m.embed.weight, pos, x, img_rows=img_rows, n_img=Lq, err_flag=m._err_flag(),
                             alias_positions=cfg.alias_embed_positions)
        # dropout: one 64-bit seed per forward (backward regenerates / re-reads this forward's masks)
        self._fw_count += 1
        rank = torch.distributed.get_rank(self.pg) if self.world > 1 else 0
        dseed = (self.seed * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03 + self._fw_count) & 0xFFFFFFFFFFFFFFFF
        pd, pa = self.p_drop, self.p_attn
        if pd > 0:                               # forward_embedding ends with dropout(x) (model.py:242-244 -> torchscale)
            ops.dropout_f32(x, p=pd, site=self.SITE_X0, seed=dseed)
        tabs = dp._xpos(T, x.device)
        scale = (D // H) ** -0.5
        ctx = dict(B=B, T=T, M=M, tabs=tabs, scale=scale, dseed=dseed)
        saved = []
        for li in range(len(self.layers)):
            # recompute: every layer writes into ONE shared set of activation buffers and only its input survives; backward
            # re-runs the layer first (the dropout masks are functions of (seed, site), so the second pass draws the same ones)
            s = self._layer_forward(li, x, ctx, tag="rc" if self.recompute else str(li))
            saved.append(dict(x_in=x) if self.recompute else s)
            x = s["x_out"]
        hF = self._buf("hF", (M, D), bf)
        ops.layernorm(x, dp.layer_norm.weight, dp.layer_norm.bias, hF, eps=cfg.eps)
        logits = self._buf("logits", (M, V), f32)
        ops.gemm(hF, self._w16(m.output_projection.weight), logits, bias=m.output_projection.bias)
        return dict(saved=saved, x_last=x, hF=hF, logits=logits, B=B, T=T, M=M, tabs=tabs, scale=scale, vis=vis, dseed=dseed)

    def _layer_forward(self, li, x, ctx, tag):
        """Training forward of decoder layer li from its fp32 input x; returns the activations backward needs (+ x_out).
        `tag` names the buffer set: one per layer normally, one shared set under activation recompute."""
        cfg, L = self.cfg, self.layers[li]
        B, T, M, tabs, scale, dseed = ctx["B"], ctx["T"], ctx["M"], ctx["tabs"], ctx["scale"], ctx["dseed"]
        D, F, H = cfg.dim, cfg.ffn, cfg.heads
        bf, f32 = torch.bfloat16, torch.float32
        pd, pa = self.p_drop, self.p_attn
        s = dict(x_in=x)
        s["h1"] = self._buf(f"h1_{tag}", (M, D), bf)
        s["qkv"] = self._buf(f"qkv_{tag}", (M, 3 * D), bf)
        s["att"] = self._buf(f"att_{tag}", (M, D), bf)
        s["lse"] = self._buf(f"lse_{tag}", (H, B, ops.lse_pad(T)), f32)
        s["a_ln"] = self._buf(f"aln_{tag}", (M, D), bf)
        s["x_mid"] = self._buf(f"xmid_{tag}", (M, D), f32)
        s["h2"] = self._buf(f"h2_{tag}", (M, D), bf)
        s["u"] = self._buf(f"u_{tag}", (M, F), bf)
        s["g_ln"] = self._buf(f"gln_{tag}", (M, F), bf)
        x_out = self._buf(f"xout_{li}", (M, D), f32)          # always per layer: it IS the next layer's saved input
        wqkv, _ = self._qkv(L, "weight")
        bqkv, _ = self._qkv(L, "bias")
        ops.layernorm(x, L["ln_a"].weight, L["ln_a"].bias, s["h1"], eps=cfg.eps)
        ops.gemm(s["h1"], wqkv, s["qkv"], bias=bqkv, xpos=tuple(tabs), seq_len=T)
        qkv = s["qkv"]
        row_mask = None
        if pa > 0:       # keep bits drawn ahead of the flash kernel: row-major for it (one buffer, reused by every layer),
                         # key-major for this layer's backward kernel (kept until then: 1 bit per score)
            words = ops.attn_dropout_mask_words(B, H, T)
            row_mask = self._buf("dmask_rows", (words,), torch.int32)
            s["dmask"] = self._buf(f"dmask_{tag}", (words,), torch.int32)
            ops.attn_dropout_masks(row_mask, s["dmask"], p=pa, site=li * 4 + 2, seed=dseed, batch=B, heads=H, seq_len=T)
        ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], s["att"], batch=B, heads=H, seq_len=T, causal=True,
                      scale=scale, lse_out=s["lse"], drop_p=pa, row_mask=row_mask)
        ops.layernorm(s["att"], L["ln_i"].weight, L["ln_i"].bias, s["a_ln"], eps=cfg.eps)
        ops.gemm(s["a_ln"], self._w16(L["o"].weight), s["x_mid"], bias=L["o"].bias, res=x, drop=(pd, li * 4, dseed))
        ops.layernorm(s["x_mid"], L["ln_f"].weight, L["ln_f"].bias, s["h2"], eps=cfg.eps)
        ops.gemm(s["h2"], self._w16(L["fc1"].weight), s["u"], bias=L["fc1"].bias)
        ops.act_layernorm(s["u"], L["ln_ffn"].weight, L["ln_ffn"].bias, s["g_ln"], eps=cfg.eps)
        ops.gemm(s["g_ln"], self._w16(L["fc2"].weight), x_out, bias=L["fc2"].bias, res=s["x_mid"], drop=(pd, li * 4 + 1, dseed))
        s["x_out"] = x_out
        return s

    # ------------------------------------------------------------------ last CLIP encoder layer (optional fine-tuning)
    def _clip_last_forward(self, x_in, N):
        """CLIPEncoderLayer.forward ([HF] modeling_clip.py:380-410) of the LAST ViT layer on the trainer's bf16 weight copies,
        keeping what backward needs: x_mid = x + out_proj(attn(qkv(LN1 x))), xv = x_mid + fc2(act(fc1(LN2 x_mid))).
        x_in: the stream after the frozen layers, fp32 [N*Tv, Dv] (a workspace of the model: copied)."""
        cfg, L = self.cfg, self.clip_last
        Tv, Dv, Hv, Fm = cfg.vit_tokens, cfg.vit_dim, cfg.vit_heads, cfg.vit_mlp
        M = N * Tv
        bf, f32 = torch.bfloat16, torch.float32
        s = dict(N=N)
        s["x_in"] = self._buf("c_xin", (M, Dv), f32)
        s["x_in"].copy_(x_in)
        for name, shape, dt in (("h1", (M, Dv), bf), ("qkv", (M, 3 * Dv), bf), ("att", (M, Dv), bf), ("x_mid", (M, Dv), f32),
                                ("h2", (M, Dv), bf), ("u", (M, Fm), bf), ("mid", (M, Fm), bf), ("xv", (M, Dv), f32),
                                ("lse", (Hv, N, ops.lse_pad(Tv)), f32)):
            s[name] = self._buf("c_" + name, shape, dt)
        act = _abi.KX_ACT_GELU if cfg.vit_act == "gelu" else _abi.KX_ACT_QUICK_GELU
        wqkv, _ = self._qkv(L, "weight")
        bqkv, _ = self._qkv(L, "bias")
        ops.layernorm(s["x_in"], L["ln1"].weight, L["ln1"].bias, s["h1"], eps=cfg.eps)
        ops.gemm(s["h1"], wqkv, s["qkv"], bias=bqkv)
        qkv = s["qkv"]
        ops.attention(qkv[:, :Dv], qkv[:, Dv:2 * Dv], qkv[:, 2 * Dv:], s["att"], batch=N, heads=Hv, seq_len=Tv, causal=False,
                      scale=(Dv // Hv) ** -0.5, lse_out=s["lse"])
        ops.gemm(s["att"], self._w16(L["o"].weight), s["x_mid"], bias=L["o"].bias, res=s["x_in"])
        ops.layernorm(s["x_mid"], L["ln2"].weight, L["ln2"].bias, s["h2"], eps=cfg.eps)
        ops.gemm(s["h2"], self._w16(L["fc1"].weight), s["u"], bias=L["fc1"].bias)
        ops.gelu_fwd(s["u"], s["mid"], act)
        ops.gemm(s["mid"], self._w16(L["fc2"].weight), s["xv"], bias=L["fc2"].bias, res=s["x_mid"])
        return s

    def _clip_last_backward(self, s, dxv, dxvb):
        """Parameter gradients of the last ViT layer from dxv = d(loss)/d(ViT output) (fp32) and its bf16 copy dxvb; nothing
        flows further (the layers below are frozen)."""
        cfg, L = self.cfg, self.clip_last
        Tv, Dv, Hv, Fm = cfg.vit_tokens, cfg.vit_dim, cfg.vit_heads, cfg.vit_mlp
        N = s["N"]
        M = N * Tv
        bf, f32 = torch.bfloat16, torch.float32
        act = _abi.KX_ACT_GELU if cfg.vit_act == "gelu" else _abi.KX_ACT_QUICK_GELU
        dmid = self._buf("c_dmid", (M, Fm), bf)
        du = self._buf("c_du", (M, Fm), bf)
        dh = self._buf("c_dh", (M, Dv), bf)
        datt = self._buf("c_datt", (M, Dv), bf)
        dqkv = self._buf("c_dqkv", (M, 3 * Dv), bf)
        dq_acc = self._buf("c_dqacc", (M, Dv), f32)
        delta = self._buf("c_delta", (Hv, N, ops.lse_pad(Tv), 2), f32)
        part = self._buf("c_part", (3, ops.ln_bwd_partials(M), Dv), f32)
        # MLP: xv = x_mid + fc2(act(fc1(LN2(x_mid))))
        ops.colsum(dxvb, self._g(L["fc2"].bias))
        ops.gemm(dxvb, self._w16(L["fc2"].weight), dmid, b_trans=True)
        ops.gemm(dxvb, s["mid"], self._g(L["fc2"].weight), a_trans=True, b_trans=True)
        ops.gelu_bwd(s["u"], dmid, du, act)
        ops.colsum(du, self._g(L["fc1"].bias))
        ops.gemm(du, self._w16(L["fc1"].weight), dh, b_trans=True)
        ops.gemm(du, s["h2"], self._g(L["fc1"].weight), a_trans=True, b_trans=True)
        ops.layernorm_bwd(s["x_mid"], dh, L["ln2"].weight, dxv, self._g(L["ln2"].weight), self._g(L["ln2"].bias), part,
                          eps=cfg.eps, dres=dxv, dxb=dxvb, d_colsum=self._g(L["o"].bias))
        # attention: x_mid = x_in + out_proj(attn(qkv(LN1(x_in))))
        ops.gemm(dxvb, self._w16(L["o"].weight), datt, b_trans=True)
        ops.gemm(dxvb, s["att"], self._g(L["o"].weight), a_trans=True, b_trans=True)
        qkv = s["qkv"]
        ops.attention_bwd(qkv[:, :Dv], qkv[:, Dv:2 * Dv], qkv[:, 2 * Dv:], s["att"], datt, s["lse"], dqkv[:, :Dv],
                          dqkv[:, Dv:2 * Dv], dqkv[:, 2 * Dv:], dq_acc, delta, batch=N, heads=Hv, seq_len=Tv, causal=False,
                          scale=(Dv // Hv) ** -0.5)
        wqkv, gwqkv = self._qkv(L, "weight")
        _, gbqkv = self._qkv(L, "bias")
        ops.colsum(dqkv, gbqkv)
        ops.gemm(dqkv, wqkv, dh, b_trans=True)
        ops.gemm(dqkv, s["h1"], gwqkv, a_trans=True, b_trans=True)
        ops.layernorm_bwd(s["x_in"], dh, L["ln1"].weight, dxv, self._g(L["ln1"].weight), self._g(L["ln1"].bias), part, eps=cfg.eps)

    # ------------------------------------------------------------------ perceiver resampler + image_proj (trainable)
    def _resampler_forward(self, xv, B, x0, T, img_rows, pos):
        """PerceiverResampler + image_proj (model.py:232) on the trainer's bf16 weight copies, keeping
        what the backward needs.  xv: ViT output, media-major [m*B*Tv, Dv] fp32 (frozen: no gradient flows into it)."""
        m, cfg = self.model, self.cfg
        Tv, Dv, Lq, Hp = cfg.vit_tokens, cfg.vit_dim, cfg.p_latents, cfg.p_heads
        inner, Fv = Hp * cfg.p_dim_head, cfg.vit_dim * cfg.p_ff_mult
        nm = len(img_rows)
        N = B * nm
        bf, f32 = torch.bfloat16, torch.float32
        pv = m.perceive
        mp = pv.media_pos_emb.view(-1, Dv)[0:nm]
        lat = self._buf("v_lat0", (N * Lq, Dv), f32)
        ops.broadcast_rows(pv.latents, lat, N)
        saved = []
        for li, (attn, ff) in enumerate(pv.layers):
            s = dict(lat_in=lat)
            s["cat"] = self._buf(f"v_cat{li}", (N * (Tv + Lq), Dv), bf)
            s["lnl"] = self._buf(f"v_lnl{li}", (N * Lq, Dv), bf)
            s["q"] = self._buf(f"v_q{li}", (N * Lq, inner), bf)
            s["kv"] = self._buf(f"v_kv{li}", (N * (Tv + Lq), 2 * inner), bf)
            s["att"] = self._buf(f"v_att{li}", (N * Lq, inner), bf)
            s["lat_mid"] = self._buf(f"v_latmid{li}", (N * Lq, Dv), f32)
            s["h"] = self._buf(f"v_h{li}", (N * Lq, Dv), bf)
            s["u"] = self._buf(f"v_u{li}", (N * Lq, Fv), bf)
            s["mid"] = self._buf(f"v_mid{li}", (N * Lq, Fv), bf)
            lat_out = self._buf(f"v_latout{li}", (N * Lq, Dv), f32)
            ops.layernorm(xv, attn.norm_media.weight, attn.norm_media.bias, s["cat"], pre_add=mp,
                          pre_add_group=Tv * B if nm > 1 else 0, grp=(Tv, Tv + Lq, 0))
            ops.layernorm(lat, attn.norm_latents.weight, attn.norm_latents.bias, s["cat"], grp=(Lq, Tv + Lq, Tv))
            ops.layernorm(lat, attn.norm_latents.weight, attn.norm_latents.bias, s["lnl"])
            ops.gemm(s["lnl"], self._w16(attn.to_q.weight), s["q"])
            ops.gemm(s["cat"], self._w16(attn.to_kv.weight), s["kv"])
            ops.perceiver_attention(s["q"], s["kv"], s["att"], batch=N, heads=Hp, n_q=Lq, n_kv=Tv + Lq, v_col_off=inner,
                                    scale=cfg.p_dim_head ** -0.5)
            ops.gemm(s["att"], self._w16(attn.to_out.weight), s["lat_mid"], res=lat)
            ops.layernorm(s["lat_mid"], ff[0].weight, ff[0].bias, s["h"])
            ops.gemm(s["h"], self._w16(ff[1].weight), s["u"])
            ops.gelu_fwd(s["u"], s["mid"])
            ops.gemm(s["mid"], self._w16(ff[3].weight), lat_out, res=s["lat_mid"])
            saved.append(s)
            lat = lat_out
        out_n = self._buf("v_outn", (N * Lq, Dv), bf)
        ops.layernorm(lat, pv.norm.weight, pv.norm.bias, out_n)
        wip = self._w16(m.image_proj.weight)
        for i, r0 in enumerate(img_rows):                # image i of all sequences: rows [i*B*64, (i+1)*B*64)
            ops.gemm(out_n[i * B * Lq:(i + 1) * B * Lq], wip, x0, grp=(Lq, T, r0), add_tab=pos, add_off=r0 + 2)
        return dict(saved=saved, lat_last=lat, out_n=out_n, xv=xv, N=N, nm=nm)

    def _resampler_backward(self, vs, dx, B, T, img_rows):
        """Gradients of image_proj and the resampler from dx = d(loss)/d(decoder input) (its image rows)."""
        m, cfg = self.model, self.cfg
        Tv, Dv, Lq, Hp, D = cfg.vit_tokens, cfg.vit_dim, cfg.p_latents, cfg.p_heads, cfg.dim
        inner, Fv = Hp * cfg.p_dim_head, cfg.vit_dim * cfg.p_ff_mult
        N, nm = vs["N"], vs["nm"]
        bf, f32 = torch.bfloat16, torch.float32
        pv = m.perceive
        R = N * Lq
        d_rows = self._buf("v_drows", (R, D), bf)
        for i, r0 in enumerate(img_rows):
            ops.gather_rows(dx, d_rows[i * B * Lq:(i + 1) * B * Lq], grp=(Lq, T, r0))
        ops.gemm(d_rows, vs["out_n"], self._g(m.image_proj.weight), a_trans=True, b_trans=True)
        dh = self._buf("v_dh", (R, Dv), bf)
        ops.gemm(d_rows, self._w16(m.image_proj.weight), dh, b_trans=True)
        dlat = self._buf("v_dlat", (R, Dv), f32)
        dlatb = self._buf("v_dlatb", (R, Dv), bf)
        part = self._buf("v_part", (3, ops.ln_bwd_partials(R), Dv), f32)
        ops.layernorm_bwd(vs["lat_last"], dh, pv.norm.weight, dlat, self._g(pv.norm.weight), self._g(pv.norm.bias), part, dxb=dlatb)
        dmid = self._buf("v_dmid", (R, Fv), bf)
        du = self._buf("v_du", (R, Fv), bf)
        datt = self._buf("v_datt", (R, inner), bf)
        dq = self._buf("v_dq", (R, inner), bf)
        dkv = self._buf("v_dkv", (N * (Tv + Lq), 2 * inner), bf)
        dcat = self._buf("v_dcat", (N * (Tv + Lq), Dv), bf)
        dxn = self._buf("v_dxn", (N * Tv, Dv), bf)
        clip = vs.get("clip")
        # d(loss)/d(ViT output): thrown away while CLIP is frozen (one block of scratch), summed over the resampler's layers
        # when its last layer trains
        dxm = self._buf("v_dxm", (N * Tv if clip else B * Tv, Dv), f32)
        dxmb = self._buf("v_dxmb", (N * Tv, Dv), bf) if clip else None
        n_pl = len(pv.layers)
        part_m = self._buf("v_partm", (3, ops.ln_bwd_partials(B * Tv), Dv), f32)
        g_mp = self._g(pv.media_pos_emb).view(-1, Dv)
        mp = pv.media_pos_emb.view(-1, Dv)
        for li in range(len(pv.layers) - 1, -1, -1):
            attn, ff = pv.layers[li]
            s = vs["saved"][li]
            # feed-forward: lat_out = lat_mid + W2 gelu(W1 LN(lat_mid))
            ops.gemm(dlatb, self._w16(ff[3].weight), dmid, b_trans=True)
            ops.gemm(dlatb, s["mid"], self._g(ff[3].weight), a_trans=True, b_trans=True)
            ops.gelu_bwd(s["u"], dmid, du)
            ops.gemm(du, self._w16(ff[1].weight), dh, b_trans=True)
            ops.gemm(du, s["h"], self._g(ff[1].weight), a_trans=True, b_trans=True)
            ops.layernorm_bwd(s["lat_mid"], dh, ff[0].weight, dlat, self._g(ff[0].weight), self._g(ff[0].bias), part, dres=dlat, dxb=dlatb)
            # cross-attention: lat_mid = lat_in + to_out(attn(to_q(LN_l(lat_in)), to_kv([LN_m(x + mp) | LN_l(lat_in)])))
            ops.gemm(dlatb, self._w16(attn.to_out.weight), datt, b_trans=True)
            ops.gemm(dlatb, s["att"], self._g(attn.to_out.weight), a_trans=True, b_trans=True)
            ops.perceiver_attention_bwd(s["q"], s["kv"], s["att"], datt, dq, dkv, batch=N, heads=Hp, n_q=Lq, n_kv=Tv + Lq,
                                        v_col_off=inner, scale=cfg.p_dim_head ** -0.5)
            ops.gemm(dq, s["lnl"], self._g(attn.to_q.weight), a_trans=True, b_trans=True)
            ops.gemm(dq, self._w16(attn.to_q.weight), dh, b_trans=True)
            ops.gemm(dkv, s["cat"], self._g(attn.to_kv.weight), a_trans=True, b_trans=True)
            ops.gemm(dkv, self._w16(attn.to_kv.weight), dcat, b_trans=True)
            ops.gather_rows(dcat, dh, grp=(Lq, Tv + Lq, Tv), accumulate=True)        # latents enter through q AND through k|v
            ops.layernorm_bwd(s["lat_in"], dh, attn.norm_latents.weight, dlat, self._g(attn.norm_latents.weight),
                              self._g(attn.norm_latents.bias), part, dres=dlat, dxb=dlatb)
            # media side: only parameter gradients (norm_media, media_pos_emb[i]); the ViT output itself is frozen
            ops.gather_rows(dcat, dxn, grp=(Tv, Tv + Lq, 0))
            for i in range(nm):
                blk = slice(i * B * Tv, (i + 1) * B * Tv)
                dst = dxm[blk] if clip else dxm
                # d_colsum sums dx INCLUDING dres: with the running sum in dst, media_pos_emb[i]'s gradient (the sum over the
                # layers of each one's column sums) is the column sum of the last call's total
                ops.layernorm_bwd(vs["xv"][blk], dxn[blk], attn.norm_media.weight, dst, self._g(attn.norm_media.weight),
                                  self._g(attn.norm_media.bias), part_m, pre_add=mp[i], accumulate=True,
                                  d_colsum=g_mp[i] if not clip or li == 0 else None,
                                  dres=dst if clip and li < n_pl - 1 else None, dxb=dxmb[blk] if clip and li == 0 else None)
        ops.sum_rows_f32(dlat.view(N, Lq * Dv), self._g(pv.latents).view(-1))
        if clip:
            self._clip_last_backward(clip, dxm, dxmb)

    # ------------------------------------------------------------------ loss + backward
    def _backward(self, fw, text_tokens, img_rows, dlogits_in=None, accumulate=False):
        """dlogits_in = None: the fused cross-entropy (loss + its gradient) starts the backward pass; otherwise the
        caller's gradient w.r.t. the logits (B, T, V) does (the autograd bridge of ``Kosmos.forward`` in train mode).
        accumulate: add this pass's (all-reduced) gradient to what the flat buffer already holds (gradient accumulation
        over micro-batches, the reference's GRADIENT_ACCUMULATE_EVERY, train.py:55,492) instead of replacing it; the kernels
        write rather than accumulate, so the previous sum is parked in a second buffer for the duration of the pass."""
        m, cfg = self.model, self.cfg
        cap = self.bwd_max_ctas if self.world > 1 else 0

        def gemm(*a_, **k_):
            return ops.gemm(*a_, max_ctas=cap, **k_)
        g_prev = None
        if accumulate:
            g_prev = self._buf("g_prev", (self.n_total,), torch.float32)
            g_prev.copy_(self.G)
        B, T, M = fw["B"], fw["T"], fw["M"]
        D, F, H, V = cfg.dim, cfg.ffn, cfg.heads, cfg.vocab
        bf, f32 = torch.bfloat16, torch.float32
        Lq = cfg.p_latents
        dp = m.decoder
        Vp = _round_up(V)
        dlogits = self._buf("dlogits", (M, Vp), bf)
        self.scalars.zero_()
        self.G.zero_()
        if dlogits_in is None:
            # targets and the number of rows that carry a loss are derived on the device (no host arithmetic on shapes,
            # pad masking included): scalars[5] = count, read by the cross-entropy kernel for its 1/count
            targets = self._buf("targets", (B, T), torch.int64)
            ops.loss_targets(text_tokens, targets, self.scalars[5:6], img_rows=img_rows, n_img=Lq, rule=self.loss_rule,
                             ignore_token=self.pad_token_id)
            ops.ce_fwd_bwd(fw["logits"], targets.view(-1), self.scalars[0:2], count=self.scalars[5:6], dlogits=dlogits,
                           err_flag=m._err_flag())
        else:       # the caller's loss lives in PyTorch: one cast of its gradient into the bf16 GEMM operand (rows padded to Vp)
            if tuple(dlogits_in.shape) != (B, T, V) or not dlogits_in.is_cuda:
                raise ValueError(f"gradient w.r.t. the logits must be a CUDA tensor of shape {(B, T, V)}")
            dlogits[:, :V].copy_(dlogits_in.reshape(M, V))
            if Vp > V:
                dlogits[:, V:].zero_()
        dl = dlogits[:, :V]
        P = ops.ln_bwd_partials(M)
        part_d = self._buf("part_d", (3, P, D), f32)
        part_f = self._buf("part_f", (3, P, F), f32)
        dx = self._buf("dx", (M, D), f32)
        dxb = self._buf("dxb", (M, D), bf)
        dh = self._buf("dh", (M, D), bf)            # gradient w.r.t. a LayerNorm output of width D (dgrad result)
        dgl = self._buf("dgl", (M, F), bf)
        du = self._buf("du", (M, F), bf)
        datt = self._buf("datt", (M, D), bf)
        dqkv = self._buf("dqkv", (M, 3 * D), bf)
        dq_acc = self._buf("dq_acc", (M, D), f32)
        delta = self._buf("delta", (H, B, ops.lse_pad(T), 2), f32)
        # LM head
        wout = m.output_projection.weight
        gemm(dl, self._w16(wout), dh, b_trans=True)
        gemm(dl, fw["hF"], self._g(wout), a_trans=True, b_trans=True)
        if m.output_projection.bias is not None:
            ops.colsum(dl, self._g(m.output_projection.bias))
        last = self.layers[-1] if self.layers else None
        pd, pa, dseed = self.p_drop, self.p_attn, fw["dseed"]
        nl = len(self.layers)
        ops.layernorm_bwd(fw["x_last"], dh, dp.layer_norm.weight, dx, self._g(dp.layer_norm.weight), self._g(dp.layer_norm.bias),
                          part_d, eps=cfg.eps, dxb=dxb, d_colsum=self._g(last["fc2"].bias) if last else None,
                          drop=(pd, (nl - 1) * 4 + 1, dseed) if last else None)
        works = []
        self._bucket_ready("head", works)                     # the LM head gradient is complete
        for li in range(len(self.layers) - 1, -1, -1):
            L, s = self.layers[li], fw["saved"][li]
            if self.recompute:               # activation checkpointing (train.py:84-110): rebuild the layer's activations from its input
                s = self._layer_forward(li, s["x_in"], fw, tag="rc")
            # ---- FFN: x_out = x_mid + fc2(LN_ffn(gelu(fc1(LN_f(x_mid)))))
            gemm(dxb, self._w16(L["fc2"].weight), dgl, b_trans=True)
            gemm(dxb, s["g_ln"], self._g(L["fc2"].weight), a_trans=True, b_trans=True)
            ops.layernorm_bwd(s["u"], dgl, L["ln_ffn"].weight, du, self._g(L["ln_ffn"].weight), self._g(L["ln_ffn"].bias), part_f,
                              act=_abi.KX_ACT_GELU, eps=cfg.eps, d_colsum=self._g(L["fc1"].bias))
            gemm(du, self._w16(L["fc1"].weight), dh, b_trans=True)
            gemm(du, s["h2"], self._g(L["fc1"].weight), a_trans=True, b_trans=True)
            ops.layernorm_bwd(s["x_mid"], dh, L["ln_f"].weight, dx, self._g(L["ln_f"].weight), self._g(L["ln_f"].bias), part_d,
                              eps=cfg.eps, dres=dx, dxb=dxb, d_colsum=self._g(L["o"].bias), drop=(pd, li * 4, dseed))
            # ---- attention: x_mid = x_in + out_proj(LN_i(attn(xpos(qkv(LN_a(x_in))))))
            gemm(dxb, self._w16(L["o"].weight), dh, b_trans=True)
            gemm(dxb, s["a_ln"], self._g(L["o"].weight), a_trans=True, b_trans=True)
            ops.layernorm_bwd(s["att"], dh, L["ln_i"].weight, datt, self._g(L["ln_i"].weight), self._g(L["ln_i"].bias), part_d,
                              eps=cfg.eps)
            qkv = s["qkv"]
            ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], s["att"], datt, s["lse"], dqkv[:, :D], dqkv[:, D:2 * D],
                              dqkv[:, 2 * D:], dq_acc, delta, batch=B, heads=H, seq_len=T, causal=True, scale=fw["scale"],
                              xpos=tuple(fw["tabs"]), drop_p=pa, drop_mask=s.get("dmask"))
            wqkv, gwqkv = self._qkv(L, "weight")
            _, gbqkv = self._qkv(L, "bias")
            ops.colsum(dqkv, gbqkv)
            gemm(dqkv, wqkv, dh, b_trans=True)
            gemm(dqkv, s["h1"], gwqkv, a_trans=True, b_trans=True)
            prev = self.layers[li - 1] if li > 0 else None
            ops.layernorm_bwd(s["x_in"], dh, L["ln_a"].weight, dx, self._g(L["ln_a"].weight), self._g(L["ln_a"].bias), part_d,
                              eps=cfg.eps, dres=dx, dxb=dxb, d_colsum=self._g(prev["fc2"].bias) if prev else None,
                              drop=(pd, (li - 1) * 4 + 1, dseed) if prev else None)
            self._bucket_ready(li, works)                     # (fc2.bias of layer li was written by layer li+1's LayerNorm backward)
        if pd > 0:                               # backward of the decoder-input dropout: the same mask on the gradient
            ops.dropout_f32(dx, p=pd, site=self.SITE_X0, seed=dseed)
        ops.embed_bwd(dx, text_tokens, self._g(m.embed.weight), self._g(m.embed_positions.weight), img_rows=img_rows, n_img=Lq,
                      padding_idx=m.embed.padding_idx if m.embed.padding_idx is not None else -1,
                      alias_positions=cfg.alias_embed_positions)
        if fw["vis"] is not None:
            self._resampler_backward(fw["vis"], dx, B, T, img_rows)
        self._bucket_ready("tail", works)
        self._finish_reduce(works)
        if dlogits_in is not None and self.world > 1:
            # autograd bridge: an external optimizer reads param.grad directly, so the all-reduced SUM becomes the MEAN over
            # ranks here (DistributedDataParallel's convention).  KosmosTrainer.step keeps the sum and folds 1/world into
            # the clip coefficient instead (one pass less over the buffer).
            self.G.mul_(1.0 / self.world)
        if g_prev is not None:
            self.G.add_(g_prev)

    # ------------------------------------------------------------------ gradient all-reduce (data parallel)
    def bucket_plan(self):
        """[(name, lo, hi)] slices of the flat gradient buffer in the order backward completes them: the LM head, then
        the layers from last to first (decay block + no-decay block each), then the tail (final LayerNorm, embedding
        and position tables).  The slices tile [0, n_total) exactly."""
        nl = len(self.layers)
        plan = [("head", self._layer_span(nl, True)[0], self.vision_decay_off)]
        for li in range(nl - 1, -1, -1):
            plan.append((f"layer{li}.decay", *self._layer_span(li, True)))
            plan.append((f"layer{li}.nodecay", *self._layer_span(li, False)))
        plan.append(("tail", self.vision_decay_off, self.n_decay))              # resampler + image_proj weights
        plan.append(("tail", self._layer_span(nl, False)[0], self.n_total))     # final LN, tables, resampler vectors
        return [(n, lo, hi) for n, lo, hi in plan if hi > lo]

    def _bucket_ready(self, which, works):
        """Overlapped all-reduce (sum): as soon as a bucket of bucket_plan() can no longer change, it goes out on
        NCCL's stream while backward continues on the compute stream.  which = "head" | layer index | "tail".
        With grad_reduce_dtype = bf16 the bucket is first cast into the bf16 exchange buffer (kx_cast_f32_to_bf16 on the
        compute stream); ``_finish_reduce`` converts the received sums back."""
        if self.world == 1 or self._defer_reduce:
            return
        if not self.overlap:
            if which == "tail":
                self._reduce_slice(0, self.n_total, works)
            return
        key = f"layer{which}." if isinstance(which, int) else which
        for name, lo, hi in self.bucket_plan():
            if name == key or (isinstance(which, int) and name.startswith(key)):
                self._reduce_slice(lo, hi, works)

    def _reduce_slice(self, lo, hi, works):
        dist = torch.distributed
        if self.grad_reduce_dtype == torch.float32:
            works.append((dist.all_reduce(self.G[lo:hi], group=self.pg, async_op=True), lo, hi))
            return
        g16 = self._buf("g16", (self.n_total,), torch.bfloat16)
        ops.cast_bf16(self.G[lo:hi], g16[lo:hi])
        works.append((dist.all_reduce(g16[lo:hi], group=self.pg, async_op=True), lo, hi))

    def _finish_reduce(self, works):
        for w, lo, hi in works:
            w.wait()
        if works and self.grad_reduce_dtype == torch.bfloat16:
            g16 = self._buf("g16", (self.n_total,), torch.bfloat16)
            ops.cast_f32(g16, self.G)               # every slice of the plan was exchanged: one pass over the whole buffer

    def _layer_span(self, li, decay):
        """[lo, hi) of layer li in the decay / no-decay segment (li == len(layers): the start of what follows them)."""
        nl = len(self.layers)
        if decay:
            first = self.seg[id(self.layers[0]["q"].weight)].off if nl else 0
            per = (self.seg[id(self.layers[1]["q"].weight)].off - first) if nl > 1 else \
                  (self.seg[id(self.model.output_projection.weight)].off - first)
        else:
            first = self.seg[id(self.layers[0]["q"].bias)].off if nl else self.n_decay
            per = (self.seg[id(self.layers[1]["q"].bias)].off - first) if nl > 1 else \
                  (self.seg[id(self.model.decoder.layer_norm.weight)].off - first)
        return first + li * per, first + (li + 1) * per

    # ------------------------------------------------------------------ optimizer
    def _optimize(self, micro_batches: int = 1):
        """Clip + optimizer over the flat buffers; the gradient is the SUM over ranks and micro-batches: scale by their count."""
        self.t += 1
        sc = self.scalars
        lr = self.lr * (float(self.lr_schedule(self.t)) if self.lr_schedule is not None else 1.0)
        self.last_lr = lr
        ops.sumsq(self.G, sc[2:3])
        ops.clip_scale(sc[2:3], self.max_grad_norm, 1.0 / (self.world * max(int(micro_batches), 1)), sc[3:4], sc[4:5])
        nd = self.n_decay
        segs = ((0, nd, self.wd, self.W16), (nd, self.n_total, 0.0, None))
        for lo, hi, wd, wb in segs:
            if hi <= lo:
                continue
            if self.opt == "adamw":
                ops.adamw_step(self.P[lo:hi], self.G[lo:hi], self.M1[lo:hi], self.M2[lo:hi], wb, lr=lr, betas=self.betas,
                               eps=self.eps, weight_decay=wd, step=self.t, grad_scale=sc[3:4])
            else:
                ops.lion_step(self.P[lo:hi], self.G[lo:hi], self.M1[lo:hi], wb, lr=lr, betas=self.betas, weight_decay=wd,
                              grad_scale=sc[3:4])
        self._inference_copies_stale()

    def _optimize_sharded(self, micro_batches: int = 1):
        """shard_optimizer: exchange + clip + optimizer (see the constructor).  self.G holds this rank's LOCAL gradient sum."""
        dist = torch.distributed
        self.t += 1
        sc = self.scalars
        lr = self.lr * (float(self.lr_schedule(self.t)) if self.lr_schedule is not None else 1.0)
        self.last_lr = lr
        nd, nt, bf16 = self.n_decay, self.n_total, torch.bfloat16
        rank = dist.get_rank(self.pg)
        chunk = self._nd_pad // self.world
        lo, hi = self.shard_range(rank)
        g16d = self._ws.get("g16d")
        if g16d is None:                         # (zeros: the padding behind the decay segment is exchanged too)
            g16d = self._ws["g16d"] = torch.zeros(self._nd_pad, dtype=bf16, device=self.P.device)
        g16s = self._buf("g16s", (chunk,), bf16)
        g16n = self._buf("g16n", (nt - nd,), bf16)
        ops.cast_bf16(self.G[:nd], g16d[:nd])
        ops.cast_bf16(self.G[nd:], g16n)
        dist.reduce_scatter_tensor(g16s, g16d, group=self.pg)
        dist.all_reduce(g16n, group=self.pg)
        ops.cast_f32(g16n, self.G[nd:])
        sc[6:8].zero_()                          # (kx_sumsq accumulates; [5] is the cross-entropy's row count)
        if hi > lo:
            ops.cast_f32(g16s[:hi - lo], self.G[lo:hi])
            ops.sumsq(self.G[lo:hi], sc[6:7])
        dist.all_reduce(sc[6:7], group=self.pg)                  # every rank receives the same bits: replicas stay identical
        ops.sumsq(self.G[nd:], sc[7:8])
        torch.add(sc[6:7], sc[7:8], out=sc[2:3])
        ops.clip_scale(sc[2:3], self.max_grad_norm, 1.0 / (self.world * max(int(micro_batches), 1)), sc[3:4], sc[4:5])
        for a, b, wd, wb in ((lo, hi, self.wd, self.W16), (nd, nt, 0.0, None)):
            if b <= a:
                continue
            w16 = wb[a:b] if wb is not None else None
            if self.opt == "adamw":
                ops.adamw_step(self.P[a:b], self.G[a:b], self.M1[a:b], self.M2[a:b], w16, lr=lr, betas=self.betas,
                               eps=self.eps, weight_decay=wd, step=self.t, grad_scale=sc[3:4])
            else:
                ops.lion_step(self.P[a:b], self.G[a:b], self.M1[a:b], w16, lr=lr, betas=self.betas, weight_decay=wd,
                              grad_scale=sc[3:4])
        dist.all_gather_into_tensor(self._W16p, self._W16p[rank * chunk:(rank + 1) * chunk], group=self.pg)
        self._masters_sharded = True
        self._inference_copies_stale()

    # ------------------------------------------------------------------ public API
    def _prepare(self, text_tokens, images, image_positions):
        cfg = self.cfg
        if not isinstance(text_tokens, torch.Tensor) or not isinstance(images, torch.Tensor):
            raise TypeError("text_tokens and images must be instances of torch.Tensor")
        if not text_tokens.is_cuda or not images.is_cuda:
            raise RuntimeError("KosmosTrainer: inputs must be on the GPU (no CPU path)")
        if text_tokens.dtype != torch.int64 or text_tokens.ndim != 2:
            raise TypeError("text_tokens must be an int64 tensor of shape (B, T_text)")
        mcount = images.shape[1] if images.ndim == 5 else 1
        pos = [2] if image_positions is None else [int(p) for p in image_positions]
        if len(pos) != mcount or sorted(pos) != pos or pos[0] < 0 or pos[-1] > text_tokens.shape[1]:
            raise ValueError(f"image_positions {pos} must be {mcount} ascending text-token indices")
        T = text_tokens.shape[1] + cfg.p_latents * mcount
        if T + 2 > cfg.max_positions:
            raise ValueError(f"spliced sequence length {T} exceeds the positional table (max {cfg.max_positions - 2})")
        img_rows = tuple(p + i * cfg.p_latents for i, p in enumerate(pos))
        images = images.to(torch.float32).reshape(-1, 3, cfg.image, cfg.image).contiguous()
        return text_tokens.contiguous(), images, img_rows

    def attach_grads(self):
        """Point every trained ``param.grad`` at its slice of the flat gradient buffer again (``optimizer.zero_grad()``
        with ``set_to_none=True``, torch's default, drops the views)."""
        for p in self.params:
            g = self._g(p)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    def autograd_forward(self, text_tokens, images, img_rows):
        """Training forward whose result is part of a PyTorch autograd graph: ``logits.backward(...)`` (through any loss
        written in PyTorch) runs the hand-scheduled backward and leaves the gradients in ``param.grad``.  See
        ``Kosmos.forward``."""
        self.sync_weights()                      # an external optimizer updates the fp32 masters only
        return _KosmosAutograd.apply(self.model.output_projection.weight, self, text_tokens, images, img_rows)

    def loss_and_grads(self, text_tokens, images, image_positions=None, accumulate: bool = False):
        """Forward + backward (+ all-reduce) without the optimizer: fills ``param.grad`` (views of the flat buffer,
        SUMMED over ranks; with ``accumulate`` also added to what was there) and returns the local mean loss (device scalar)."""
        text_tokens, images, img_rows = self._prepare(text_tokens, images, image_positions)
        fw = self._forward(text_tokens, images, img_rows)
        self._backward(fw, text_tokens, img_rows, accumulate=accumulate)
        self._last_fw = fw                       # (tests read the dropout seed / recorded attention masks of the step)
        return self.scalars[0] / torch.clamp(self.scalars[1], min=1.0)

    def step_accumulated(self, micro_batches):
        """One optimisation step over several micro-batches (gradient accumulation, train.py:55,492): ``micro_batches`` is a
        sequence of ``(text_tokens, images)`` or ``(text_tokens, images, image_positions)``; the update uses the mean of
        their gradients.  Returns the mean of the micro-batch losses (device scalar)."""
        micro_batches = list(micro_batches)
        if not micro_batches:
            raise ValueError("step_accumulated needs at least one micro-batch")
        total = None
        self._defer_reduce = self.shard_optimizer        # sharded optimizer: ONE exchange, of the accumulated local sums
        try:
            for i, mb in enumerate(micro_batches):
                loss = self.loss_and_grads(*mb, accumulate=i > 0).clone()
                total = loss if total is None else total + loss
        finally:
            self._defer_reduce = False
        if self.shard_optimizer:
            self._optimize_sharded(len(micro_batches))
        else:
            self._optimize(len(micro_batches))
        return total / len(micro_batches)

    def step(self, text_tokens, images, image_positions=None):
        self._defer_reduce = self.shard_optimizer
        try:
            loss = self.loss_and_grads(text_tokens, images, image_positions)
        finally:
            self._defer_reduce = False
        if self.shard_optimizer:
            self._optimize_sharded()
        else:
            self._optimize()
        return loss

    @property
    def grad_norm(self):
        """Global gradient norm of the last step before clipping (device scalar)."""
        return self.scalars[4]


def cosine_with_warmup(warmup_steps: int, total_steps: int, num_cycles: float = 0.5):
    """LR multiplier of the reference's default schedule (train.py:206-251 -> transformers ``get_cosine_schedule_with_warmup``,
    selected at train.py:560-583 with 1 % warm-up): linear 0 -> 1 over ``warmup_steps``, then half a cosine to 0 at
    ``total_steps``.  The optimizer step k (1-based) uses the multiplier of scheduler step k - 1, as ``optim.step();
    scheduler.step()`` does (train.py:655-656).  Pass as ``KosmosTrainer(lr_schedule=...)``."""
    import math

    def mult(step: int) -> float:
        cur = step - 1
        if cur < warmup_steps:
            return cur / max(1, warmup_steps)
        progress = (cur - warmup_steps) / max(1, total_steps - warmup_steps)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * num_cycles * 2.0 * progress)))
    return mult


def linear_with_warmup(warmup_steps: int, total_steps: int):
    """train.py:206-251 with ``scheduler_type="linear"``: transformers ``get_linear_schedule_with_warmup``."""
    def mult(step: int) -> float:
        cur = step - 1
        if cur < warmup_steps:
            return cur / max(1, warmup_steps)
        return max(0.0, (total_steps - cur) / max(1, total_steps - warmup_steps))
    return mult


class _KosmosAutograd(torch.autograd.Function):
    """Bridge between PyTorch autograd and the hand-scheduled backward: the graph has ONE node for the whole model.
    ``anchor`` (a trained parameter) only makes the output require grad; the parameter gradients are written where the
    trainer always writes them (the flat buffer that every ``param.grad`` views), so the node returns no gradients."""

    @staticmethod
    def forward(ctx, anchor, trainer, text_tokens, images, img_rows):
        fw = trainer._forward(text_tokens, images, img_rows)
        trainer._fw_serial += 1
        ctx.trainer, ctx.fw, ctx.text_tokens, ctx.img_rows, ctx.serial = trainer, fw, text_tokens, img_rows, trainer._fw_serial
        return fw["logits"].view(fw["B"], fw["T"], -1)

    @staticmethod
    def backward(ctx, dlogits):
        tr = ctx.trainer
        if ctx.fw is None or ctx.serial != tr._fw_serial:
            raise RuntimeError("kosmosx: the activations of this forward are gone (a later forward re-used the buffers, or "
                               "backward ran twice); call backward once, before the next training forward")
        # autograd semantics: add to the gradients that are there; after zero_grad(set_to_none=True) nothing is there
        accumulate = all(p.grad is not None and p.grad.data_ptr() == tr._g(p).data_ptr() for p in tr.params)
        tr._backward(ctx.fw, ctx.text_tokens, ctx.img_rows, dlogits_in=dlogits, accumulate=accumulate)
        ctx.fw = None
        tr.attach_grads()
        return None, None, None, None, None
