"""Incremental-decoding benchmark: Kosmos.generate at the reference size, random-init weights.
Reports the prompt pass, the time per one-token step (CUDA events around the graph replays) and the achieved HBM
bandwidth of a step = (decoder weights streamed once + KV cache read + new KV rows written) / step time, against the
measured copy bandwidth in MEASURED_PEAKS.json.  Usage: python tools/bench_decode.py [--batch 8] [--prompt 512] [--new 128]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "kosmos-x_b200"))
from kosmosx import Kosmos, KosmosConfig, ops  # noqa: E402


def decode_bytes(cfg, batch, n_cached):
    """Algorithmic HBM bytes of ONE decoding step with n_cached tokens already in the cache."""
    d, f, L, v = cfg.dim, cfg.ffn, cfg.layers, cfg.vocab
    weights = 2 * (L * (4 * d * d + 2 * d * f) + v * d)             # bf16, every matrix read once
    kv = 2 * L * batch * (n_cached + 1) * 2 * d                     # bf16 k and v rows 0..n_cached
    return weights, kv


def run(model, batch, prompt, new, layers_note="", one_kernel=True):
    cfg = model.cfg
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    text = torch.randint(0, cfg.vocab, (batch, prompt - cfg.p_latents), generator=g).to(dev)
    images = torch.randn(batch, 3, cfg.image, cfg.image, generator=g).to(dev)
    model.generate(text, images, 4)                                 # stages weights, warms the allocator
    torch.cuda.synchronize()
    e0, ep, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    # time the pieces of generate() separately: prompt pass, then `new - 1` graph replays
    text_p, images_p, img_rows, T = model._prepare_inputs(text, images, None, new)
    dp = model.decoder._pack()
    x0 = model._ws.get("x0", (batch * T, cfg.dim), torch.float32, dev)
    e0.record()
    xv = model._vit(images_p, media=1)
    model._perceive_project(xv, batch, x0, T, img_rows)
    ops.embed_splice_pos(text_p, dp["embed"], dp["pos"], x0, img_rows=img_rows, n_img=cfg.p_latents,
                         alias_positions=cfg.alias_embed_positions)
    dec = model.decoder
    state, _ = dec.begin_generation(x0, batch, T, T + new, head="last")
    history = torch.zeros(batch, new, dtype=torch.int64, device=dev)
    dec.advance(state, history=history, move=False)
    ep.record()
    torch.cuda.synchronize()
    trace = None
    if one_kernel:
        n_ph = 1 + 5 * cfg.layers + 2
        trace = torch.zeros(2 * n_ph + 1 + 16, dtype=torch.int64, device=dev) if os.environ.get("KX_STEP_TRACE") else None
        if trace is not None:
            trace[2 * n_ph] = int(os.environ.get("KX_STEP_TRACE_PHASE", "-1"))
        plan = dec.build_step_plan(state, history=history, trace=trace)

        class _One:
            @staticmethod
            def replay():
                ops.decode_step(plan)
        graph, nodes = _One, 1
    else:
        graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(graph):
            dec.decode_step(state)
            dec.advance(state, history=history)
        nodes = ops.launch_count() - n0
    for _ in range(3):                                              # warm-up steps (advance the position too)
        graph.replay()
    torch.cuda.synchronize()
    first_pos = int(state.pos.item())
    steps = new - 1 - 3
    e1.record()
    for _ in range(steps):
        graph.replay()
    e2.record()
    torch.cuda.synchronize()
    ms_step = e1.elapsed_time(e2) / steps
    if trace is not None:                                           # per-phase timeline of CTA 0 in the LAST step
        n_ph = 1 + 5 * cfg.layers + 2
        fine = trace[2 * n_ph + 1:].cpu().tolist()
        if fine[0]:
            lab = ["enter", "a issued", "ring 0 full", "k loop done", "sync1", "sync2", "epilogue done", "-", "barrier enter",
                   "arrived", "barrier left", "phase start", "is linear", "is nt2"]
            print("fine trace (us since item entry): " + ", ".join(f"{lab[i]}={(fine[i] - fine[0]) / 1e3:.2f}" for i in range(14) if fine[i]),
                  file=sys.stderr)
        t = trace[:2 * n_ph].cpu().view(-1, 2).tolist()
        names = ["embed"] + ["qkv", "attn", "out", "fc1", "fc2"] * cfg.layers + ["head", "pick"]
        agg = {}
        prev_leave = None
        for i, (done, leave) in enumerate(t):
            if prev_leave is not None and done:
                a = agg.setdefault(names[i], [0, 0.0, 0.0])
                a[0] += 1; a[1] += (done - prev_leave) / 1e3
                if leave:
                    a[2] += (leave - done) / 1e3
            prev_leave = leave if leave else None
        if os.environ.get("KX_STEP_TRACE_RAW"):
            t0 = t[0][0]
            for i in range(56, 72):
                print(f"raw {i:3d} {names[i]:5s} done={(t[i][0] - t0) / 1e3:9.2f} leave={(t[i][1] - t0) / 1e3:9.2f}", file=sys.stderr)
        for k, (n, work, wait) in agg.items():
            print(f"trace {k:6s} x{n:3d}: CTA0 work {work / n:7.2f} us, barrier wait {wait / n:7.2f} us", file=sys.stderr)
    ms_prompt = e0.elapsed_time(ep)
    last_pos = int(state.pos.item())
    w_bytes, kv_lo = decode_bytes(cfg, batch, first_pos)
    _, kv_hi = decode_bytes(cfg, batch, last_pos - 1)
    by = w_bytes + 0.5 * (kv_lo + kv_hi)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6530.6))
    out = dict(batch=batch, prompt=T, new_tokens=new, timed_steps=steps, kernels_per_step=nodes, ms_prompt_pass=ms_prompt, ms_per_step=ms_step,
               tokens_per_s=batch / ms_step * 1e3, weight_bytes=w_bytes, kv_bytes_mean=0.5 * (kv_lo + kv_hi),
               achieved_gbs=by / ms_step / 1e6, peak_gbs=peak, frac=by / ms_step / 1e6 / peak,
               path=(f"one persistent kernel per step ({int(ops.lib.kx_decode_step_ctas())} CTAs)" if one_kernel
                     else "per-kernel step, CUDA-graph replay"),
               err_flag=int(state.err.item()), tokens_head=history[0, :8].tolist())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--prompt", type=int, default=512)
    ap.add_argument("--new", type=int, default=128)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--per-kernel", action="store_true", help="also time every launch of one eager step")
    ap.add_argument("--mode", default="one", choices=["one", "graph"], help="one persistent kernel per step, or per-kernel graph")
    a = ap.parse_args()
    cfg = KosmosConfig(max_positions=2050, layers=a.layers)
    torch.manual_seed(0)
    model = Kosmos(config=cfg, device="cuda")
    r = run(model, a.batch, a.prompt, a.new, one_kernel=(a.mode == "one"))
    print(json.dumps(r))
    if a.per_kernel:
        st = model._last_decode_state
        ops.profile_begin()
        model.decoder.decode_step(st)
        recs = ops.profile_end()
        agg = {}
        for kind, fl, by, ms in recs:
            k = agg.setdefault(kind, [0, 0.0, 0.0])
            k[0] += 1; k[1] += ms; k[2] += by
        for kind, (n, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{kind:32s} x{n:3d}  {ms / n * 1e3:8.1f} us each  {by / max(ms, 1e-9) / 1e6:8.0f} GB/s")


if __name__ == "__main__":
    main()
