"""world_size-2 gloo tests (CPU) of the N>1 host logic: batch sharding, barrier, max-over-ranks
timing and logits reassembly.  The forward has no data-path collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kosmosx import dist as kd


def test_shard_range_partitions_exactly():
    for gb in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [kd.shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        kd.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = kd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    gb = 5
    text = torch.arange(gb * 3).view(gb, 3)
    images = torch.arange(gb * 2, dtype=torch.float32).view(gb, 2)
    t, im = kd.shard_batch(text, images, r, w)
    # stand-in for the per-rank forward: a function of the local shard only
    local = (t.float().sum(1, keepdim=True) + im.sum(1, keepdim=True)).view(-1, 1, 1)
    kd.barrier()
    full = kd.gather_logits(local, gb)
    want = (text.float().sum(1, keepdim=True) + images.sum(1, keepdim=True)).view(-1, 1, 1)
    ok = torch.equal(full, want)
    mx = kd.max_over_ranks(10.0 + rank)
    sm = kd.sum_over_ranks(t.shape[0])
    out[rank] = (ok, mx, sm)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reductions():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True, 11.0, 5.0), 1: (True, 11.0, 5.0)}


# --------------------------------------------------------------------------- training step: gradient buckets
def _tiny_trainer_layout():
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig, KosmosTrainer
    oc = ko.OracleConfig.tiny(layers=3)
    torch.manual_seed(0)
    model = Kosmos(config=KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__}))
    return model, KosmosTrainer(model, layout_only=True)


def test_gradient_bucket_plan_tiles_the_flat_buffer():
    """ONE all-reduce over the trainable gradients per step, issued as buckets in the order backward
    completes them (LM head, layers last to first, tables).  The buckets must tile the flat buffer exactly, keep every
    parameter inside one bucket, and hold exactly the trained set (.A branches, no .B, no vision tower)."""
    model, tr = _tiny_trainer_layout()
    plan = tr.bucket_plan()
    spans = sorted((lo, hi) for _, lo, hi in plan)
    assert spans[0][0] == 0 and spans[-1][1] == tr.n_total
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    names = [n for n, _, _ in plan]
    assert names[0] == "head" and names[-2:] == ["tail", "tail"] and names[1].startswith("layer2.") and names[-3].startswith("layer0.")
    trained = {n for n, p in model.named_parameters() if id(p) in {id(q) for q in tr.params}}
    assert {"perceive.latents", "perceive.media_pos_emb", "image_proj.weight", "perceive.layers.0.0.to_kv.weight"} <= trained
    by_id = {id(p): n for n, p in model.named_parameters()}
    for p in tr.params:
        s = tr.seg[id(p)]
        inside = [n for n, lo, hi in plan if lo <= s.off and s.off + s.numel <= hi]
        assert len(inside) == 1, by_id[id(p)]
        name = by_id[id(p)]
        assert ".B." not in name and not name.startswith("clip_model")       # CLIP tower frozen, multiway .B branches never run
        if name.startswith("decoder.layers."):
            assert inside[0].startswith(f"layer{name.split('.')[2]}.")
    # q|k|v are adjacent so that one [3D, D] GEMM operand / gradient view exists
    L = tr.layers[1]
    sq, sk, sv = (tr.seg[id(L[n].weight)] for n in "qkv")
    assert sk.off == sq.off + sq.numel and sv.off == sk.off + sk.numel
    # decay split of the reference (train.py:257-398): Linear weights decay, biases / LayerNorm / embeddings do not
    for p in tr.params:
        assert (tr.seg[id(p)].off < tr.n_decay) == (p.ndim == 2 and by_id[id(p)].endswith("weight")
                                                    and "embed" not in by_id[id(p)])


def test_bucket_plan_with_the_last_clip_layer_unfrozen():
    """train_clip_last_layer=True (notes.txt:537-538): the last ViT layer's 16 tensors join the tail buckets (its gradient is
    the last one backward produces), q|k|v adjacent, Linear weights in the decay segment; the plan still tiles the buffer."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig, KosmosTrainer
    oc = ko.OracleConfig.tiny(layers=3)
    model = Kosmos(config=KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__}))
    base = KosmosTrainer(model, layout_only=True)
    tr = KosmosTrainer(model, layout_only=True, train_clip_last_layer=True)
    assert len(tr.params) == len(base.params) + 16
    plan = tr.bucket_plan()
    spans = sorted((lo, hi) for _, lo, hi in plan)
    assert spans[0][0] == 0 and spans[-1][1] == tr.n_total
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    by_id = {id(p): n for n, p in model.named_parameters()}
    last = f"clip_model.encoder.layers.{oc.vit_layers - 1}."
    clip = [p for p in tr.params if by_id[id(p)].startswith("clip_model")]
    assert len(clip) == 16 and all(by_id[id(p)].startswith(last) for p in clip)
    for p in clip:
        s = tr.seg[id(p)]
        assert [n for n, lo, hi in plan if lo <= s.off and s.off + s.numel <= hi] == ["tail"]
        assert (s.off < tr.n_decay) == (p.ndim == 2)
    a = model.clip_model.encoder.layers[-1].self_attn
    sq, sk, sv = (tr.seg[id(x.weight)] for x in (a.q_proj, a.k_proj, a.v_proj))
    assert sk.off == sq.off + sq.numel and sv.off == sk.off + sk.numel
    with pytest.raises(ValueError):
        KosmosTrainer(model, layout_only=True, train_clip_last_layer=True, train_resampler=False)


def _bucket_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    kd.init_from_env("gloo")
    _, tr = _tiny_trainer_layout()
    tr.world, tr.overlap = world, True
    tr.grad_reduce_dtype = torch.float32          # (the bf16 exchange casts with a CUDA kernel; its NCCL test is tools/dp_check.py)
    g = torch.Generator().manual_seed(100 + rank)
    tr.G = torch.randn(tr.n_total, generator=g)
    want = tr.G.clone()
    dist.all_reduce(want)
    works = []
    tr._bucket_ready("head", works)
    for li in range(len(tr.layers) - 1, -1, -1):
        tr._bucket_ready(li, works)
    tr._bucket_ready("tail", works)
    tr._finish_reduce(works)
    ok_overlap = torch.equal(tr.G, want) and len(works) == len(tr.bucket_plan())
    tr.G = torch.randn(tr.n_total, generator=torch.Generator().manual_seed(100 + rank))
    tr.overlap = False
    works = []
    tr._bucket_ready("head", works)
    tr._bucket_ready(0, works)
    tr._bucket_ready("tail", works)
    tr._finish_reduce(works)
    out[rank] = (ok_overlap, torch.equal(tr.G, want), len(works))
    dist.destroy_process_group()


def test_two_rank_gloo_bucketed_gradient_all_reduce():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True, True, 1), 1: (True, True, 1)}


# --------------------------------------------------------------------------- sharded optimizer (ZeRO-1): host logic
class _TorchOps:
    """torch stand-ins for the CUDA kernels the optimizer path calls — this test is about the exchange / slicing logic of
    KosmosTrainer._optimize_sharded (which collective moves which slice), not about the kernels (tools/dp_check.py runs the
    real thing on NCCL)."""

    @staticmethod
    def cast_bf16(src, dst=None):
        if dst is None:
            return src.to(torch.bfloat16)
        dst.copy_(src)
        return dst

    @staticmethod
    def cast_f32(src, dst):
        dst.copy_(src)
        return dst

    @staticmethod
    def sumsq(g, out):
        out += (g.double() ** 2).sum().float()

    @staticmethod
    def clip_scale(sumsq_t, max_norm, pre_scale, scale_out, norm_out=None):
        norm = sumsq_t.sqrt() * pre_scale
        if norm_out is not None:
            norm_out.copy_(norm)
        c = torch.clamp(max_norm / (norm + 1e-6), max=1.0) if max_norm > 0 else torch.ones(1)
        scale_out.copy_(c * pre_scale)

    @staticmethod
    def adamw_step(p, g, m, v, wb, *, lr, betas, eps, weight_decay, step, grad_scale=None):
        gi = g * (grad_scale if grad_scale is not None else 1.0)
        p.mul_(1.0 - lr * weight_decay)
        m.mul_(betas[0]).add_(gi, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(gi, gi, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        p.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))
        if wb is not None:
            wb.copy_(p)


def _sharded_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    kd.init_from_env("gloo")
    import kosmos_oracle as ko
    import kosmosx.train as ktrain
    from kosmosx import Kosmos, KosmosConfig, KosmosTrainer
    ktrain.ops = _TorchOps
    oc = ko.OracleConfig.tiny(layers=3)

    def make(shard):
        torch.manual_seed(0)
        model = Kosmos(config=KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__}))
        tr = KosmosTrainer(model, layout_only=True, shard_optimizer=shard, lr=1e-2, weight_decay=0.1)
        n = tr.n_total
        g = torch.Generator().manual_seed(7)
        tr.P = torch.randn(n, generator=g)
        tr.M1, tr.M2 = torch.zeros(n), torch.zeros(n)
        tr._W16p = torch.zeros(tr._nd_pad, dtype=torch.bfloat16)
        tr.W16 = tr._W16p[:tr.n_decay]
        tr.W16.copy_(tr.P[:tr.n_decay])
        tr.scalars = torch.zeros(8)
        return tr

    a, b = make(True), make(False)
    ok = a.shard_optimizer and a.world == world and not b.shard_optimizer
    spans = [a.shard_range(r) for r in range(world)]
    ok &= spans[0][0] == 0 and spans[-1][1] == a.n_decay and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    ok &= a._nd_pad % (world * 1024) == 0 and a._nd_pad >= a.n_decay
    for step in range(3):
        g_local = torch.randn(a.n_total, generator=torch.Generator().manual_seed(100 * step + rank)) * 0.1
        a.G, b.G = g_local.clone(), g_local.clone()
        a.scalars.zero_(); b.scalars.zero_()
        a._optimize_sharded()
        works = []
        b._bucket_ready("tail", works)           # one bf16 all-reduce of the whole buffer, converted back to fp32
        b._finish_reduce(works)
        b._optimize()
    lo, hi = a.shard_range()
    nd = a.n_decay
    own = all(torch.equal(getattr(a, k)[lo:hi], getattr(b, k)[lo:hi]) and torch.equal(getattr(a, k)[nd:], getattr(b, k)[nd:])
              for k in ("P", "M1", "M2"))
    stale = not torch.equal(a.P[:nd], b.P[:nd])                  # the other rank's slice has not been updated here
    raised = False
    try:
        a.model.state_dict()                                     # (layout_only trainers register no hook: ask the guard directly)
        a._require_whole_masters("state_dict()")
    except RuntimeError:
        raised = True
    a.gather_masters()
    whole = torch.equal(a.P, b.P) and torch.equal(a.W16, b.W16) and not a._masters_sharded
    norm_same = torch.equal(a.scalars[4], b.scalars[4])
    out[rank] = (bool(ok), bool(own), bool(stale), raised, bool(whole), bool(norm_same))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_optimizer_equals_the_all_reduce_step():
    """KosmosTrainer(shard_optimizer=True): reduce-scatter of the bf16 decay-segment gradients + each rank's slice of the update +
    all-gather of the bf16 copies gives, after gather_masters(), exactly the masters / copies of the all-reduce + replicated
    optimizer (a sum of two values has one order), the replicated tail is updated everywhere, and state_dict() is refused
    while the masters are sharded."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True,) * 6, 1: (True,) * 6}


def test_sharded_optimizer_plan_and_argument_checks():
    """shard_range() cuts the padded decay segment into `world` equal, 1024-aligned chunks that tile it; a single rank or
    layout without torch.distributed keeps the replicated optimizer; the options that cannot be combined are refused."""
    _, tr = _tiny_trainer_layout()
    assert not tr.shard_optimizer and tr._nd_pad == tr.n_decay            # world == 1: nothing to shard
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig, KosmosTrainer
    oc = ko.OracleConfig.tiny(layers=3)
    model = Kosmos(config=KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__}))
    one = KosmosTrainer(model, layout_only=True, shard_optimizer=True)
    assert not one.shard_optimizer                                        # asked for, but this process trains alone
    for world in (2, 3, 8, 64):
        tr.world, tr.shard_optimizer = world, True
        tr._nd_pad = (tr.n_decay + world * 1024 - 1) // (world * 1024) * (world * 1024)
        spans = [tr.shard_range(r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == tr.n_decay
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        chunk = tr._nd_pad // world
        assert chunk % 1024 == 0 and all(lo % 1024 == 0 or lo == tr.n_decay for lo, _ in spans)
        assert all(hi - lo == chunk for lo, hi in spans if hi < tr.n_decay)   # only the last non-empty slice may be short
    tr._masters_sharded = True
    with pytest.raises(RuntimeError, match="gather_masters"):
        tr._require_whole_masters("state_dict()")
