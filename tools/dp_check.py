"""Multi-rank correctness of the data-parallel training step on NCCL hardware (reference:
train.py:709 `init_process_group("nccl")`, train.py:643-657 the step).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Every rank builds the same tiny model (seeded), takes its contiguous shard of one global batch and checks, for the bucketed
all-reduce overlapped with backward AND the single all-reduce after it, in fp32 and bf16 exchange:
  1. the all-reduced gradient / world == the gradient of ONE process on the concatenated batch (fp32 reduction-order
     tolerance), and the mean of the rank losses == its loss;
  2. after 3 optimizer steps the replicas are BIT-IDENTICAL (every rank compares its flat master / moment buffers with
     rank 0's) and within Adam's sign-flip bound of the single-process run;
  3. the autograd bridge (model.train(); loss.backward(); torch.optim) leaves the MEAN gradient in param.grad;
  4. KosmosTrainer(shard_optimizer=True) (reduce-scatter, each rank updates its slice, all-gather of the bf16 copies) reproduces
     the all-reduce run after 3 steps — bit for bit at 2 ranks — and state_dict() refuses to run until gather_masters().
Exit code 0 = all checks passed on every rank.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kosmos-x_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig, KosmosTrainer
    from kosmosx import dist as kd
    rank, local, world = kd.init_from_env("nccl")
    torch.cuda.set_device(local)
    oc = ko.OracleConfig.tiny(max_positions=512)
    kc = KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__})
    sd = ko.build(oc, seed=0).state_dict()
    per = 2
    text, images = ko.make_inputs(oc, per * world, 40, seed=5)
    lo, hi = kd.shard_range(per * world, rank, world)
    assert hi - lo == per
    tg, ig = text.cuda(), images.cuda()
    failures = []

    def check(ok, what):
        if not ok:
            failures.append(what)
            print(f"[rank {rank}] FAIL {what}", flush=True)

    def fresh(**kw):
        m = Kosmos(config=kc)
        m.load_state_dict(sd)
        m = m.cuda()
        return m, KosmosTrainer(m, lr=1e-3, weight_decay=0.1, dropout=0.0, attention_dropout=0.0, **kw)

    # the single-process answer on the concatenated batch
    m1, t1 = fresh(distributed=False)
    loss1 = t1.loss_and_grads(tg, ig).item()
    g1 = t1.G.clone()
    for _ in range(3):
        t1.step(tg, ig)
    p1 = t1.P.clone()
    for overlap in (True, False):
        for rd in (torch.float32, torch.bfloat16):
            tag = f"overlap={overlap} reduce={str(rd).split('.')[-1]}"
            m2, t2 = fresh(overlap_all_reduce=overlap, grad_reduce_dtype=rd)
            check(t2.world == world, f"{tag}: trainer sees world {t2.world}")
            loss = t2.loss_and_grads(tg[lo:hi], ig[lo:hi])
            lsum = loss.clone()
            dist.all_reduce(lsum)
            check(abs(lsum.item() / world - loss1) <= 2e-5 * abs(loss1), f"{tag}: mean of rank losses {lsum.item() / world} vs {loss1}")
            rel = ((t2.G / world - g1).norm() / g1.norm()).item()
            tol = 2e-5 if rd == torch.float32 else 6e-3            # bf16 exchange: one rounding of every summand
            check(rel <= tol, f"{tag}: all-reduced gradient vs single process rel {rel:.3e} > {tol}")
            for _ in range(3):
                t2.step(tg[lo:hi], ig[lo:hi])
            torch.cuda.synchronize()
            same = True
            for name, buf in (("P", t2.P), ("M1", t2.M1), ("M2", t2.M2), ("W16", t2.W16)):
                ref0 = buf.clone()
                dist.broadcast(ref0, 0)
                same &= bool(torch.equal(ref0, buf))
                check(torch.equal(ref0, buf), f"{tag}: replica buffer {name} differs from rank 0's")
            same_t = torch.tensor([int(same)], device="cuda")
            dist.all_reduce(same_t, op=dist.ReduceOp.MIN)
            d = (t2.P - p1).abs()
            # Adam's first steps move every element by ~lr * sign(g): an element whose gradient is at the reduction-noise
            # level may take the other sign, nothing else may differ
            check(d.max().item() <= 3 * 2.1e-3, f"{tag}: parameters {d.max().item():.3e} from the single-process run")
            frac = (d > 1e-5).float().mean().item()
            mean_d = d.mean().item()
            # (the fraction grows with the world size: more summands reordered, more near-zero gradients whose Adam step flips sign;
            # measured 1.4e-2 at 2 ranks, 6.6e-2 at 8 with mean |diff| 3e-6 — the invariants that matter are the gradient itself,
            # checked above, and the bit-identical replicas)
            check(frac <= (0.15 if rd == torch.float32 else 0.3) and mean_d <= (2e-5 if rd == torch.float32 else 2e-4),
                  f"{tag}: {frac:.3e} of the parameters differ from the single-process run (mean |diff| {mean_d:.2e})")
            if rank == 0:
                print(f"{tag}: grad rel {rel:.2e}, params max diff {d.max().item():.2e}, differing fraction {frac:.2e}, "
                      f"replicas {'bit-identical' if same_t.item() else 'DIFFER'} after 3 steps", flush=True)
            if not overlap and rd == torch.bfloat16:
                ar = {k: getattr(t2, k).clone() for k in ("P", "M1", "M2", "W16")}      # what the sharded optimizer must reproduce
            del m2, t2
    # sharded optimizer (ZeRO-1): reduce-scatter + each rank's slice of the update + all-gather of the bf16 copies
    m3, t3 = fresh(shard_optimizer=True)
    for _ in range(3):
        t3.step(tg[lo:hi], ig[lo:hi])
    try:
        m3.state_dict()
        check(False, "sharded: state_dict() before gather_masters() did not raise")
    except RuntimeError:
        pass
    slo, shi = t3.shard_range()
    own = all(torch.equal(getattr(t3, k)[slo:shi], ar[k][slo:shi]) for k in ("P", "M1", "M2"))
    nd = t3.n_decay
    tail = all(torch.equal(getattr(t3, k)[nd:], ar[k][nd:]) for k in ("P", "M1", "M2"))
    t3.gather_masters()
    m3.state_dict()
    torch.cuda.synchronize()
    dP, dW = (t3.P - ar["P"]).abs(), (t3.W16.float() - ar["W16"].float()).abs()
    if world == 2:          # a sum of two bf16 values has one order: the sharded step must reproduce the all-reduce step bit for bit
        check(own and tail, "sharded: this rank's slice of P / M1 / M2 (or the replicated tail) differs from the all-reduce run")
        check(torch.equal(t3.P, ar["P"]) and torch.equal(t3.W16, ar["W16"]), "sharded: gathered masters / bf16 copies differ from the all-reduce run")
    else:                   # (ring order of the reduce-scatter differs from the all-reduce's)
        check(dP.max().item() <= 3 * 2.1e-3 and dP.mean().item() <= 2e-4, f"sharded: masters {dP.max().item():.3e} / mean {dP.mean().item():.2e} from the all-reduce run")
    same = True
    for name, buf in (("P", t3.P), ("W16", t3.W16)):
        ref0 = buf.clone()
        dist.broadcast(ref0, 0)
        same &= bool(torch.equal(ref0, buf))
        check(torch.equal(ref0, buf), f"sharded: replica buffer {name} differs from rank 0's after gather_masters()")
    m3.eval()
    with torch.no_grad():
        lg = m3(tg[lo:hi], ig[lo:hi])          # inference after training on the gathered weights
    check(bool(torch.isfinite(lg).all()), "sharded: inference after gather_masters() is not finite")
    if rank == 0:
        print(f"shard_optimizer: slice {slo}..{shi} of {nd}; vs the all-reduce run: masters max diff {dP.max().item():.2e}, bf16 copies max diff "
              f"{dW.max().item():.2e}, own slice + tail {'bit-identical' if own and tail else 'differ'}, replicas "
              f"{'bit-identical' if same else 'DIFFER'} after gather", flush=True)
    del m3, t3
    # autograd bridge: param.grad is the MEAN over ranks (DistributedDataParallel's convention)
    tgt = ko.KosmosOracle.loss_targets(text, oc.p_latents).cuda()
    def bridge(model, tok, img, tg_):
        model.train()
        for p in model.parameters():
            p.grad = None
        logits = model(tok, img)
        torch.nn.functional.cross_entropy(logits.reshape(-1, oc.vocab), tg_.reshape(-1), ignore_index=-100).backward()
        model.eval()
    mb1, tb1 = fresh(distributed=False)
    bridge(mb1, tg, ig, tgt)
    gb1 = tb1.G.clone()
    mb2, tb2 = fresh()
    bridge(mb2, tg[lo:hi], ig[lo:hi], tgt[lo:hi])
    rel = ((tb2.G - gb1).norm() / gb1.norm()).item()
    check(rel <= 1e-2, f"bridge: param.grad vs single-process mean gradient rel {rel:.3e}")     # dlogits pass through bf16
    if rank == 0:
        print(f"autograd bridge: mean-gradient rel {rel:.2e}", flush=True)
    bad = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(bad)
    if rank == 0:
        print("DP_CHECK_OK" if bad.item() == 0 else f"DP_CHECK_FAILED ({int(bad.item())} failures)", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if bad.item() == 0 else 1)


if __name__ == "__main__":
    main()
