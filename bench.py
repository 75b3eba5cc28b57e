#!/usr/bin/env python
"""bench.py — multimodal tokens/s of the Kosmos-X forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one ``Kosmos.forward(text_tokens, images)`` over one batch of synthetic input at
BASELINE.json configs[2]: B=8 sequences per GPU, spliced sequence length 2048 (1984 text tokens +
64 image latents), one 224x224 image per sequence, random-init weights of the reference
architecture (ViT-L/14 + PerceiverResampler + 24-layer d=2048 sub-LN/xPos decoder + 32002-wide head).
The position table is built with 2050 rows because the reference's 2048-row table caps T at 2046
(SURVEY.md fact 6); everything else is the reference configuration.

Rank 0 prints ONE JSON line.  Keys beyond the base contract:
  roofline      dominant kernel (the tcgen05 GEMM): algorithmic FLOPs / CUDA-event time of its launches,
                measured in an instrumented step right after the timed region, vs MEASURED_PEAKS.json
  decoder_block BASELINE.json configs[1]: one fused decoder layer (B=8, T=2048) TFLOP/s and fraction of peak
  breakdown     per-kernel-class share of one step (ms)
  cpu_baseline  the oracle (CPU restatement of the reference path) timed on this box's host cores, N=1 only
  e2e           same metric through Kosmos.forward with pinned HOST buffers: per step H2D of tokens+images
                and D2H of the full logits, all inside the timed region
  gpu_eager_baseline  SURVEY §8(d): the restated reference path in stock eager PyTorch bf16 on the same GPU and batch
  train_step    BASELINE.json configs[3]: the data-parallel training step on the same shapes (short run)
  decode        SURVEY §8(f)2: greedy decoding against the KV cache (B=8, 512-row prompt), generated tokens/s and
                its own HBM roofline (weights + cache bytes per step vs the measured copy bandwidth); N=1 only
``--impl reference`` times the oracle on the host cores (the reference's own dependencies are not
installable offline: SURVEY.md §8(c)); it is the only other place bench.py executes oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kosmos-x_b200"))

METRIC = "multimodal tokens/sec (seq=2048, 224^2 img), Kosmos.forward"
UNIT = "tokens/s"
SEQ = 2048
N_LATENTS = 64
T_TEXT = SEQ - N_LATENTS
BATCH_PER_GPU = 8
VOCAB = 32002

# --workload: the default is the configuration the metric is quoted on (configs[2]); c5 is the multi-image stress
# (configs[4]: 4 images per sequence at spliced rows 2, 514, 1026, 1538, B=4 per GPU -> 32 global at 8 GPUs)
WORKLOADS = {
    "c3": dict(batch=8, images=1, positions=None, t_text=SEQ - N_LATENTS,
               name="configs[2]: full Kosmos forward (ViT-L/14 + perceiver + 24-layer decoder + LM head), B=8 per GPU, "
                    "seq=2048 (1984 text + 64 image latents), 1 image 224x224 per sequence"),
    "strict": dict(batch=8, images=1, positions=None, t_text=2046 - N_LATENTS, seq=2046, max_positions=2048,
                   name="configs[2] at the reference's own limit: its 2048-row position table (model.py:164) caps the spliced length at "
                        "T = 2046 (SURVEY.md fact 6): B=8 per GPU, 1982 text + 64 image latents, 1 image 224x224 per sequence"),
    "c5": dict(batch=4, images=4, positions=[2, 450, 898, 1346], t_text=SEQ - 4 * N_LATENTS,
               name="configs[4]: interleaved 4 images/seq (perceiver + image-splice stress), B=4 per GPU, seq=2048 "
                    "(1792 text + 4x64 image latents at spliced rows 2/514/1026/1538)"),
}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def _ncu_traffic(launch_type: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of this type, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json: {launch type: bytes}); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(launch_type)
    except Exception:
        return None


def forward_flops_per_seq(T=SEQ, images=1):
    """2mnk per GEMM, decoder attention causal-halved."""
    d, f, L = 2048, 8192, 24
    dec_layer = 24 * T * d * d + 2 * T * T * d
    dec = L * dec_layer + 2 * T * d * VOCAB
    vit = 162.02e9 * images
    per = 4.115e9 * images
    return dec + vit + per, dec_layer


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU reference (oracle)
def cpu_reference(steps: int, warmup: int, budget_s: float = 150.0, wl=None):
    """The reference path restated on the CPU (oracle/kosmos_oracle.py), fp32, all host threads.
    One step = ONE sequence of the benchmark workload (T=2048): 1/B of a GPU step."""
    wl = wl or WORKLOADS["c3"]
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kosmos_oracle as ko
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    SEQ = wl.get("seq", 2048)
    cfg = ko.OracleConfig(max_positions=wl.get("max_positions", SEQ + 2), multiway=False)    # .B branches never execute
    model = ko.build(cfg, seed=0)
    small = ko.make_inputs(cfg, 1, 50, seed=1)
    multi = wl["images"] > 1
    text, images = ko.make_inputs(cfg, 1, wl["t_text"], seed=1, n_images=wl["images"] if multi else None)
    kw = dict(image_positions=wl["positions"]) if multi else {}
    times = []
    t_begin = time.perf_counter()
    with torch.no_grad():
        model(*small)                                              # page in weights / thread pool
        warmup = min(warmup, 1)                                    # one full-size warm-up pass is enough on a CPU
        for i in range(warmup + steps):
            if i >= warmup + 1 and time.perf_counter() - t_begin > budget_s:
                break                                              # bounded sample: stop once the budget is spent
            t0 = time.perf_counter()
            out = model(text, images, **kw)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    assert out.shape == (1, SEQ, VOCAB)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=SEQ / (ms / 1e3), ms_per_step=ms, steps_done=len(times), cores=cores,
                sample=f"1 sequence (T={SEQ}: {wl['t_text']} text tokens + {wl['images']} image(s)) per step, fp32 eager PyTorch oracle, "
                       f"{cores} threads, {len(times)} timed step(s)")


def gpu_eager_reference(torch, wl, steps=3):
    """the restated reference path (oracle) in stock eager PyTorch, bf16, on the SAME B200 and the same
    batch — what the reference's own nn.Module stack dispatches to (cuBLAS GEMMs, eager softmax / LayerNorm / xPos
    kernels, the per-layer CPU->GPU mask upload).  A like-for-like GPU baseline next to the CPU one; checker code, timed
    here only as a baseline."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kosmos_oracle as ko
    SEQ = wl.get("seq", 2048)
    cfg = ko.OracleConfig(max_positions=wl.get("max_positions", SEQ + 2), multiway=False)
    with torch.device("cuda"):
        ref = ko.KosmosOracle(cfg)
    ref = ref.to(dtype=torch.bfloat16).eval()
    B = wl["batch"]
    multi = wl["images"] > 1
    text, images = ko.make_inputs(cfg, B, wl["t_text"], seed=1, n_images=wl["images"] if multi else None)
    text, images = text.cuda(), images.cuda().to(torch.bfloat16)
    kw = dict(image_positions=wl["positions"]) if multi else {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        out = ref(text, images, **kw)                               # warm-up (cuBLAS heuristics, allocator)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            out = ref(text, images, **kw)
        e1.record()
        torch.cuda.synchronize()
    assert out.shape == (B, SEQ, VOCAB)
    ms = e0.elapsed_time(e1) / steps
    del ref, out
    torch.cuda.empty_cache()
    return {"value": B * SEQ / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "kind": "port",
            "what": f"oracle (restated reference path) in eager PyTorch bf16 on the same GPU, B={B}, seq={SEQ}"}


def run_reference(args, rank):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    r = cpu_reference(args.steps, max(args.warmup, 0), wl=wl)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps_done"], "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"].split(":")[0] + f" sample: Kosmos.forward on 1 sequence, seq={wl.get('seq', 2048)} "
                               f"({wl['t_text']} text + {wl['images']}x64 image latents), {wl['images']} image(s) 224x224, CPU",
                   "global_batch": 1, "seq_len": wl.get("seq", 2048), "parallelism": "cpu"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference deps (torchscale+passed_x patch, flamingo_pytorch, bitsandbytes) are not installable "
                "offline; this is the oracle restatement of the reference path (kind=port)",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def _init_dist(torch, kdist):
    """Rendezvous + the first collective with fd 1 pointed at stderr: NCCL prints its version banner on stdout when the
    communicator is created, and stdout must carry exactly one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, local, world = kdist.init_from_env("nccl")
        torch.cuda.set_device(local)
        if world > 1:
            t = torch.ones(1, device=torch.device("cuda", local))
            torch.distributed.all_reduce(t)
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return rank, local, world


def run_gpu(args):
    import torch
    from kosmosx import Kosmos, KosmosConfig, ops
    from kosmosx import dist as kdist
    rank, local, world = _init_dist(torch, kdist)
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun for N>1", file=sys.stderr)
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = _peaks()
    wl = WORKLOADS[args.workload]
    B, n_img, t_text = wl["batch"], wl["images"], wl["t_text"]
    SEQ = wl.get("seq", 2048)                              # (shadows the module constant: the strict workload runs T = 2046)
    max_pos = wl.get("max_positions", SEQ + 2)
    fkw = dict(image_positions=wl["positions"]) if n_img > 1 else {}

    numa = kdist.bind_host_thread_to_gpu(local)            # pinned buffers below land on the GPU's own NUMA node
    ldt = torch.bfloat16 if args.logits == "bf16" else torch.float32
    torch.manual_seed(0)                                   # same replicated random-init weights on every rank
    # graph_alias_output: the e2e loop below double-buffers the result itself (the D2H of step i overlaps step i+1)
    model = Kosmos(config=KosmosConfig(max_positions=max_pos), device=dev, cuda_graph=args.graph, graph_alias_output=True,
                   logits_dtype=ldt)
    g = torch.Generator().manual_seed(1 + rank)            # each rank owns its shard of the global batch
    h_text = torch.randint(0, VOCAB, (B, t_text), dtype=torch.long, generator=g).pin_memory()
    h_img = torch.randn(*((B, 3, 224, 224) if n_img == 1 else (B, n_img, 3, 224, 224)), generator=g).pin_memory()
    d_text, d_img = h_text.to(dev), h_img.to(dev)

    def step_resident():
        return model(d_text, d_img, **fkw)

    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    torch.cuda.synchronize()
    assert out.shape == (B, SEQ, VOCAB)
    model.check_tokens()
    del out

    # ---- timed region 1: inputs resident in HBM ("value")
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kdist.barrier(); torch.cuda.synchronize()
    if sampler: sampler.start()
    n0 = ops.launch_count()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    torch.cuda.synchronize(); kdist.barrier()
    launches = ops.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    ms_total = kdist.max_over_ranks(e0.elapsed_time(e1), dev)
    ms_step = ms_total / args.steps
    tokens_per_step = B * SEQ * world
    value = tokens_per_step / (ms_step / 1e3)

    # ---- timed region 2: end to end through the public API with host buffers ("e2e")
    # H2D of this step's tokens+images from pinned memory, forward, D2H of the full fp32 logits into pinned
    # memory.  The D2H runs on a copy stream from one of two logits buffers so it overlaps the next forward.
    copy_stream = torch.cuda.Stream(dev)
    ld_logits = VOCAB if ldt == torch.float32 else (VOCAB + 7) // 8 * 8      # bf16 rows are padded to a 16-byte pitch
    h_out = [torch.empty(B, SEQ, ld_logits, dtype=ldt).pin_memory() for _ in range(2)]
    d_keep = [None, None]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def step_e2e(i):
        s = i & 1
        t = h_text.to(dev, non_blocking=True)
        im = h_img.to(dev, non_blocking=True)
        torch.cuda.current_stream().wait_event(done[s])        # the logits slot written now was copied out (step i-2)
        logits = model(t, im, **fkw)
        d_keep[s] = logits                                     # keep alive until its copy is done
        logits.record_stream(copy_stream)
        ready = torch.cuda.Event(); ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            # the (B, T, vocab) result is a view of rows with pitch ld_logits: copy the dense buffer (one plain memcpy)
            src = logits if ld_logits == VOCAB else logits.as_strided((B, SEQ, ld_logits), (SEQ * ld_logits, ld_logits, 1))
            h_out[s].copy_(src, non_blocking=True)
            done[s].record(copy_stream)

    for i in range(2):
        step_e2e(i)
    torch.cuda.synchronize()
    kdist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    copy_stream.synchronize()
    torch.cuda.current_stream().wait_stream(copy_stream)
    e1.record()
    torch.cuda.synchronize(); kdist.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = kdist.max_over_ranks(max(e0.elapsed_time(e1), wall_ms), dev) / args.steps
    e2e_value = tokens_per_step / (e2e_ms / 1e3)
    h2d = h_text.numel() * 8 + h_img.numel() * 4
    d2h = h_out[0].numel() * h_out[0].element_size()
    checksum = float(h_out[(args.steps - 1) & 1][0, -1, :8].float().sum())
    d_keep[:] = [None, None]

    # ---- the host-side ceiling of e2e: pinned D2H bandwidth with EVERY rank copying at once (no compute running)
    probe = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h_probe = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    h_probe.copy_(probe, non_blocking=True); torch.cuda.synchronize()
    kdist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(4):
        h_probe.copy_(probe, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(); kdist.barrier()
    d2h_gbs_rank = 4 * probe.numel() / (kdist.max_over_ranks(e0.elapsed_time(e1), dev) * 1e-3) / 1e9   # slowest rank
    del probe, h_probe
    d2h_floor_ms = d2h / (d2h_gbs_rank * 1e9) * 1e3       # the step's result alone, at that rate

    # ---- instrumented step: per-kernel CUDA-event times (roofline leg), graph off
    was_graph, model.cuda_graph = model.cuda_graph, False
    model(d_text, d_img, **fkw); torch.cuda.synchronize()
    ops.profile_begin()
    for _ in range(2):
        model(d_text, d_img, **fkw)
    recs = ops.profile_end()
    model.cuda_graph = was_graph
    agg, shapes = {}, {}
    for kind, fl, by, ms in recs:
        if kind.startswith("gemm "):
            sh = shapes.setdefault(kind[5:], [0, 0.0, 0.0, 0.0])
            sh[0] += 1; sh[1] += fl; sh[2] += by; sh[3] += ms
            kind = "gemm"
        a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += fl; a[2] += by; a[3] += ms
    tot_ms = sum(a[3] for a in agg.values())
    gemm = agg["gemm"]
    gemm_tflops = gemm[1] / (gemm[3] * 1e-3) / 1e12
    top_name, top = max(shapes.items(), key=lambda kv: kv[1][3])          # the launch type with the largest time share
    top_tflops = top[1] / (top[3] * 1e-3) / 1e12
    breakdown = {k: {"launches_per_step": a[0] // 2, "ms_per_step": a[3] / 2, "share": a[3] / tot_ms,
                     **({"tflops": a[1] / (a[3] * 1e-3) / 1e12} if a[1] else {"gbs": a[2] / (a[3] * 1e-3) / 1e9})}
                 for k, a in sorted(agg.items(), key=lambda kv: -kv[1][3])}
    gemm_shapes = {k: {"launches_per_step": a[0] // 2, "us_per_launch": 1e3 * a[3] / a[0], "tflops": a[1] / (a[3] * 1e-3) / 1e12}
                   for k, a in sorted(shapes.items(), key=lambda kv: -kv[1][3])[:8]}

    # ---- configs[1]: one decoder layer alone (B=8, T=2048)
    dec = model.decoder
    x = torch.randn(BATCH_PER_GPU * SEQ, 2048, device=dev)
    one = dec._pack()["layers"][:1]
    dec_block_ms = _time_decoder_block(torch, dec, one, x, BATCH_PER_GPU, SEQ)
    flops_seq, dec_layer_flops = forward_flops_per_seq(T=SEQ, images=n_img)
    blk_tflops = dec_layer_flops * BATCH_PER_GPU / (dec_block_ms * 1e-3) / 1e12

    step_tflops = flops_seq * B / (ms_step * 1e-3) / 1e12        # per GPU
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": wl["name"] + f", random-init weights, max_positions={max_pos}",
                   "global_batch": B * world, "seq_len": SEQ, "parallelism": f"dp{world}",
                   "l2": "no flush needed: each step streams 3.3 GB of weights and >10 GB of activations (L2 = 126 MB)",
                   "cuda_graph": bool(args.graph), "logits": args.logits},
        "step_tflops_per_gpu": step_tflops,
        "step_frac_of_bf16_peak": {"burst": step_tflops / peaks["burst"], "sustained": step_tflops / peaks["sustained"]},
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel, launch type '%s' (%d launches/step; largest time share)"
                               % (top_name, top[0] // 2),
                     "achieved": top_tflops, "peak": peaks["sustained"], "peak_burst": peaks["burst"], "unit": "TFLOP/s",
                     "frac": top_tflops / peaks["sustained"], "frac_of_burst": top_tflops / peaks["burst"],
                     "flops_per_launch": top[1] / top[0], "us_per_launch": 1e3 * top[3] / top[0],
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                     "traffic": _ncu_traffic(top_name), "share_of_step": top[3] / tot_ms,
                     "all_gemm_launches": {"launches_per_step": gemm[0] // 2, "achieved": gemm_tflops,
                                           "frac": gemm_tflops / peaks["sustained"],
                                           "frac_of_burst": gemm_tflops / peaks["burst"], "share_of_step": gemm[3] / tot_ms}},
        "gemm_launch_types": gemm_shapes,
        "decoder_block": {"config": f"configs[1]: one decoder layer, B=8, T={SEQ}, d=2048, 32 heads, bf16",
                          "ms": dec_block_ms, "tflops": blk_tflops, "frac_of_burst": blk_tflops / peaks["burst"],
                          "frac_of_sustained": blk_tflops / peaks["sustained"]},
        "breakdown": breakdown,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "result": f"full {args.logits} logits (B, 2048, 32002) copied to pinned host memory "
                "(double-buffered on a copy stream; the reference run in 16-bit returns 16-bit logits)", "checksum": checksum,
                "d2h_gbs_per_gpu_all_ranks_copying": d2h_gbs_rank, "d2h_ms_per_step_at_that_rate": d2h_floor_ms,
                "host_ceiling": {"value": tokens_per_step / (max(d2h_floor_ms, 1e-9) * 1e-3), "unit": UNIT,
                                 "what": "tokens/s if a step cost only the D2H of its logits at the measured pinned-copy rate"},
                "numa": numa},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if args.train_leg and args.workload in ("c3",):
        # configs[3] beside the headline: the data-parallel training step on the same shapes (short run)
        del x, h_out, d_keep
        model._ws.clear(); model._graphs = {}; model.decoder._ws.clear()
        torch.cuda.empty_cache()
        try:
            line["train_step"] = train_measure(torch, kdist, dev, model, wl, max(2, min(args.steps, 5)), 3, world, rank, peaks,
                                               dropout=args.dropout, reduce_bf16=bool(args.reduce_bf16))
        except Exception as e:                                  # the forward line must still be printed
            line["train_step"] = {"error": f"{type(e).__name__}: {e}"}
        x = None
    if args.decode_leg and args.workload == "c3" and world == 1:
        # SURVEY §8(f)2 beside the headline: greedy decoding against a KV cache (HBM-bound; its own roofline)
        try:
            model._ws.clear(); model._graphs = {}; model.decoder._ws.clear()
            torch.cuda.empty_cache()
            line["decode"] = decode_measure(torch, model)
        except Exception as e:
            line["decode"] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and rank == 0 and not args.no_cpu:
        try:
            model._ws.clear(); model._graphs = {}; model.decoder._ws.clear()
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_reference(torch, wl)
        except Exception as e:
            line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and rank == 0 and not args.no_cpu:
        del x
        r = cpu_reference(1, 0, wl=wl)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": r["sample"], "ms_per_sample": r["ms_per_step"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    kdist.barrier()


def decode_measure(torch, model, batch=8, prompt=512, new=96):
    """Incremental decoding (Kosmos.generate's step): B=8 sequences, 512-token multimodal prompt, one token per sequence
    per step from the KV cache.  Timed with CUDA events around the steps (default path for B <= 8: ONE persistent
    cooperative kernel of libkosmosx_sm100.so per step).  Roofline: HBM — algorithmic bytes of a step = every decoder weight matrix
    once (bf16) + the K/V rows of the cache, against the measured copy bandwidth."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_decode
    r = bench_decode.run(model, batch, prompt, new, one_kernel=True)
    return {
        "config": f"greedy decoding, B={batch}, prompt {r['prompt']} rows (1 image + text), {r['timed_steps']} timed one-token steps, "
                  "KV cache bf16 head-major, " + r["path"],
        "value": r["tokens_per_s"], "unit": "generated tokens/s", "ms_per_step": r["ms_per_step"],
        "ms_prompt_pass": r["ms_prompt_pass"], "gpu_launches_per_step": r["kernels_per_step"],
        "roofline": {"bound": "hbm", "kernel": "decode_step_kernel (the whole step: embedding, 24 x (q|k|v, attention, out_proj, fc1, fc2), LM head, greedy choice)",
                     "achieved": r["achieved_gbs"], "peak": r["peak_gbs"], "unit": "GB/s", "frac": r["frac"],
                     "bytes_per_step": r["weight_bytes"] + r["kv_bytes_mean"], "traffic": _ncu_traffic("decode_step"),
                     "peak_source": "measured copy bandwidth (MEASURED_PEAKS.json)"},
    }


def train_measure(torch, kdist, dev, model, wl, steps, warmup, world, rank, peaks, optimizer="adamw", detail=True, dropout=0.1,
                  reduce_bf16=True, clip_last=False, recompute=False):
    """BASELINE.json configs[3]: the data-parallel training step (forward that keeps activations -> CE over text rows ->
    backward -> bucketed NCCL all-reduce overlapped with backward -> clip -> fused AdamW), B sequences per GPU.
    Returns a dict: tokens/s with inputs resident, e2e with H2D of the batch and D2H of the loss every step, and
    (detail) the per-kernel-class breakdown of one instrumented step."""
    from kosmosx import KosmosTrainer, ops
    from kosmosx.train import cosine_with_warmup
    B, n_img, t_text = wl["batch"], wl["images"], wl["t_text"]
    fkw = dict(image_positions=wl["positions"]) if n_img > 1 else {}
    # the reference's training mode: dropout = attention_dropout = 0.1 (model.py:175-177), cosine schedule with warm-up
    trainer = KosmosTrainer(model, optimizer=optimizer, lr=1e-5, weight_decay=0.1, max_grad_norm=1.0, dropout=dropout,
                            attention_dropout=dropout, grad_reduce_dtype=torch.bfloat16 if reduce_bf16 else torch.float32,
                            lr_schedule=cosine_with_warmup(2, 10000), overlap_all_reduce=bool(int(os.environ.get("KX_BENCH_OVERLAP", "0"))),
                            bwd_max_ctas=int(os.environ.get("KX_BENCH_BWD_CTAS", "0")), train_clip_last_layer=clip_last,
                            recompute=recompute,
                            # N > 1: ZeRO-1 style optimizer (reduce-scatter / each rank's slice of AdamW / all-gather of the bf16
                            # copies) unless KX_BENCH_SHARD=0 asks for the all-reduce + replicated optimizer
                            shard_optimizer=(world > 1 and reduce_bf16 and not int(os.environ.get("KX_BENCH_OVERLAP", "0"))
                                             and bool(int(os.environ.get("KX_BENCH_SHARD", "1")))))
    g = torch.Generator().manual_seed(11 + rank)
    h_text = torch.randint(0, VOCAB, (B, t_text), dtype=torch.long, generator=g).pin_memory()
    h_img = torch.randn(*((B, 3, 224, 224) if n_img == 1 else (B, n_img, 3, 224, 224)), generator=g).pin_memory()
    d_text, d_img = h_text.to(dev), h_img.to(dev)
    for _ in range(max(warmup, 3)):
        loss = trainer.step(d_text, d_img, **fkw)
    torch.cuda.synchronize()
    first_loss = float(loss)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kdist.barrier(); torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0.record()
    for _ in range(steps):
        loss = trainer.step(d_text, d_img, **fkw)
    e1.record()
    torch.cuda.synchronize(); kdist.barrier()
    launches = ops.launch_count() - n0
    ms = kdist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    tokens = B * SEQ * world
    # e2e: host batch -> device, step, loss -> host, every step
    h_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    kdist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        t = h_text.to(dev, non_blocking=True)
        im = h_img.to(dev, non_blocking=True)
        h_loss.copy_(trainer.step(t, im, **fkw).view(1), non_blocking=True)
    e1.record()
    torch.cuda.synchronize(); kdist.barrier()
    wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = kdist.max_over_ranks(max(e0.elapsed_time(e1), wall), dev) / steps
    fwd_flops, _ = forward_flops_per_seq(images=n_img)
    dec_flops = fwd_flops - (162.02e9 + 4.115e9) * n_img
    step_flops = (3.0 * dec_flops + (162.02e9 + 4.115e9) * n_img) * B           # frozen vision side: forward only
    tfl = step_flops / (ms * 1e-3) / 1e12
    out = {"config": "configs[3]: data-parallel training step, B=%d per GPU, seq=2048, %d image(s)/seq, bf16 operands / fp32 "
                     "master weights, %s, grad clip 1.0, cosine LR schedule, decoder + LM head + embedding tables + perceiver resampler + "
                     "image_proj trained (%s), dropout = attention_dropout = %.2f, %s gradient all-reduce, %s"
                     % (B, n_img, optimizer, "CLIP frozen except its last encoder layer" if clip_last else "CLIP tower frozen", dropout,
                        "bf16" if reduce_bf16 else "fp32", "activation recompute per decoder layer" if recompute else "no activation recompute"),
           "value": tokens / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "global_batch": B * world,
           "step_tflops_per_gpu": tfl, "step_frac_of_bf16_peak": {"burst": tfl / peaks["burst"], "sustained": tfl / peaks["sustained"]},
           "flops_per_step_per_gpu": step_flops,
           "e2e": {"value": tokens / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": (h_text.numel() * 8 + h_img.numel() * 4) * world, "d2h_bytes_per_step": 4 * world,
                   "result": "mean loss copied to pinned host memory"},
           "gpu_launches": int(launches), "loss_first": first_loss, "loss_last": float(h_loss[0]),
           "trained_parameters": int(sum(p.numel() for p in trainer.params)),
           "grad_all_reduce": (("%d buckets (per decoder layer) on NCCL's stream, overlapped with backward" % len(trainer.bucket_plan()))
                               if trainer.overlap else
                               "sharded optimizer: reduce-scatter of the bf16 Linear-weight gradients after backward, each rank updates 1/%d "
                               "of the fp32 masters, all-gather of the bf16 copies (small replicated tail: one bf16 all-reduce)" % world
                               if trainer.shard_optimizer else
                               "one all-reduce of the flat %s gradient buffer after backward (not overlapped: "
                               "profiles/r2_nccl_overlap.md)" % ("bf16" if reduce_bf16 else "fp32")) if world > 1 else "none (1 GPU)"}
    if detail:
        ops.profile_begin()
        trainer.step(d_text, d_img, **fkw)
        recs = ops.profile_end()
        agg = {}
        for kind, fl, by, t_ms in recs:
            if kind.startswith("gemm "):
                kind = "gemm wgrad" if "+tn" in kind else "gemm dgrad" if "+nn" in kind else "gemm fwd"
            a = agg.setdefault(kind, [0, 0.0, 0.0])
            a[0] += 1; a[1] += fl; a[2] += t_ms
        tot = sum(a[2] for a in agg.values())
        out["breakdown"] = {k: {"launches": a[0], "ms": a[2], "share": a[2] / tot,
                                **({"tflops": a[1] / (a[2] * 1e-3) / 1e12} if a[1] else {})}
                            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])}
    trainer.gather_masters()                 # (collective; every rank is here) the model is used again after this leg
    del trainer
    return out


def run_train(args):
    """--workload train: the training step as the benchmark line (same contract as the forward line)."""
    import torch
    from kosmosx import Kosmos, KosmosConfig
    from kosmosx import dist as kdist
    rank, local, world = _init_dist(torch, kdist)
    args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = _peaks()
    wl = WORKLOADS["c3"]
    torch.manual_seed(0)
    model = Kosmos(config=KosmosConfig(max_positions=SEQ + 2), device=dev)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler: sampler.start()
    r = train_measure(torch, kdist, dev, model, wl, args.steps, args.warmup, world, rank, peaks, optimizer=args.optimizer,
                      dropout=args.dropout, reduce_bf16=bool(args.reduce_bf16), clip_last=bool(args.clip_last),
                      recompute=bool(args.recompute))
    clocks = sampler.stop() if sampler else None
    bd = r.get("breakdown", {})
    top = max(((k, v) for k, v in bd.items() if k.startswith("gemm")), key=lambda kv: kv[1]["ms"], default=(None, None))
    line = {"metric": "multimodal tokens/sec (seq=2048, 224^2 img), training step", "value": r["value"], "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": r["config"], "global_batch": r["global_batch"], "seq_len": SEQ, "parallelism": f"dp{world}",
                       "l2": "no flush needed: each step streams > 50 GB (L2 = 126 MB)"},
            "step_tflops_per_gpu": r["step_tflops_per_gpu"], "step_frac_of_bf16_peak": r["step_frac_of_bf16_peak"],
            "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": clocks, "breakdown": bd,
            "loss_first": r["loss_first"], "loss_last": r["loss_last"], "trained_parameters": r["trained_parameters"],
            "grad_all_reduce": r["grad_all_reduce"]}
    if top[0] is not None:
        line["roofline"] = {"bound": "tensor", "kernel": f"gemm_bf16_kernel, launch class '{top[0]}' ({top[1]['launches']} launches/step)",
                            "achieved": top[1]["tflops"], "peak": peaks["sustained"], "peak_burst": peaks["burst"], "unit": "TFLOP/s",
                            "frac": top[1]["tflops"] / peaks["sustained"], "frac_of_burst": top[1]["tflops"] / peaks["burst"],
                            "peak_source": peaks["source"] + ", sustained figure", "traffic": None, "share_of_step": top[1]["share"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    kdist.barrier()


def _time_decoder_block(torch, dec, one_layer, x, B, SEQ, iters=10):
    """configs[1]: the per-layer launch sequence of Decoder.run_layers on one layer."""
    packed = dec._packed
    saved = packed["layers"]
    packed["layers"] = one_layer
    try:
        run = lambda: dec.run_layers(x, B, SEQ, head=False)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    finally:
        packed["layers"] = saved


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--graph", type=int, default=1, help="replay the forward as one CUDA graph (default on)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--logits", default="bf16", choices=["bf16", "fp32"], help="dtype of the logits Kosmos.forward returns "
                    "(bf16: the LM head's TMA-store epilogue writes 16-bit rows, as the reference run in 16-bit does; fp32: 2.1 GB per "
                    "GPU per step, the round-1 setting)")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["train"], help="c3 = configs[2] (the metric's "
                    "configuration, default); strict = the same at the reference's position-table limit T = 2046; c5 = configs[4], "
                    "4 images per sequence; train = configs[3], the training step")
    ap.add_argument("--optimizer", default="adamw", choices=["adamw", "lion"])
    ap.add_argument("--dropout", type=float, default=0.1, help="dropout = attention_dropout of the training step (reference: 0.1)")
    ap.add_argument("--clip-last", type=int, default=0, help="--workload train: also train CLIP's last encoder layer (notes.txt:537-538)")
    ap.add_argument("--recompute", type=int, default=0, help="--workload train: activation recompute per decoder layer (train.py:84-110)")
    ap.add_argument("--reduce-bf16", type=int, default=1, help="exchange gradients in bf16 (default) or fp32 (0)")
    ap.add_argument("--decode-leg", type=int, default=1, help="also time incremental decoding (SURVEY 8(f)2) and report it as "
                    "'decode' inside the forward line (default on, single GPU)")
    ap.add_argument("--train-leg", type=int, default=1, help="also time a few training steps (configs[3]) and report them "
                    "as 'train_step' inside the forward line (default on)")
    args = ap.parse_args()
    if args.steps < 1:
        ap.error("--steps must be >= 1")
    if args.impl == "reference":
        if args.workload == "train":
            args.workload = "c3"
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    if args.workload == "train":
        run_train(args)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
