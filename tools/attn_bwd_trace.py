"""Timeline of one CTA of the attention-backward kernel (kx_attn_bwd_set_trace): clock64 stamps of compute thread 0
and of the MMA-issuing thread of one CTA (default CTA 0 = batch 0, head 0, key block 0: 16 query blocks at T = 2048), and a
per-SM occupancy summary from the {SM id, entry, first score tile, exit} records every CTA leaves.
    python tools/attn_bwd_trace.py [cta ...] > gpurun_out/attn_bwd_trace.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kosmos-x_b200"))
from kosmosx import _abi, ops  # noqa: E402

dev = torch.device("cuda")
B, H, T = 8, 32, 2048
D, M = H * 64, B * T
NB = (T + 127) // 128
CTAS = B * H * NB
qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
out = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(H, B, ops.lse_pad(T), device=dev)
ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, batch=B, heads=H, seq_len=T, causal=True, scale=0.125, lse_out=lse)
d_out = torch.randn(M, D, device=dev).bfloat16()
dqkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
acc = torch.empty(M, D, device=dev)
delta = torch.empty(*lse.shape, 2, device=dev)
run = lambda: ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, d_out, lse, dqkv[:, :D], dqkv[:, D:2 * D],
                                dqkv[:, 2 * D:], acc, delta, batch=B, heads=H, seq_len=T, causal=True, scale=0.125)
for _ in range(2):
    run()
print("compute thread 0 points: 0 iter start, 1 S^T ready, 2 dP^T ready, 4 P^T / dS^T computed, 5 previous dQ MMA done (smem dS^T, P^T free),"
      " 6 P^T / dS^T stored + arrived")
print("MMA thread points: 0 loop top, 1 next S^T + dP^T issued, 2 P^T / dS^T seen, 3 dV + dK issued, 4 previous dQ read out, 5 dQ issued")
for cta in [int(a) for a in sys.argv[1:]] or [0]:
    buf = torch.full((1024 + 4 * CTAS,), -1, dtype=torch.int64, device=dev)
    buf[1023] = cta
    _abi.check(_abi.lib.kx_attn_bwd_set_trace(buf.data_ptr()), "kx_attn_bwd_set_trace")
    run()
    torch.cuda.synchronize()
    _abi.lib.kx_attn_bwd_set_trace(None)
    full = buf.cpu()
    full[1023] = -1
    t = full[:1024].view(2, 32, 16)
    t0 = int(t[t >= 0].min())
    n_it = NB - cta % NB
    print(f"=== CTA {cta} (key block {cta % NB}, {n_it} query blocks)")
    for role, name in ((0, "compute"), (1, "MMA")):
        print("---", name)
        for it in range(n_it):
            row = t[role, it]
            print(f"it {it:2d}: " + " ".join(f"{int(v) - t0:7d}" if v >= 0 else "     -1" for v in row[:7 if role == 0 else 6]))
    rec = full[1024:].view(CTAS, 4)
    g0 = int(rec[:, 1].min())
    span = int(rec[:, 3].max()) - g0
    busy, first, gaps = {}, {}, []
    for sm in rec[:, 0].unique().tolist():
        r = rec[rec[:, 0] == sm]
        r = r[r[:, 1].argsort()]
        busy[sm] = int((r[:, 3] - r[:, 1]).sum())
        first[sm] = int((r[:, 2] - r[:, 1]).sum())
        gaps += (r[1:, 1] - r[:-1, 3]).tolist()
    nb = len(busy)
    gaps = torch.tensor(gaps, dtype=torch.float64)
    own = rec[cta]
    print(f"--- per-CTA records (globaltimer): kernel span {span / 1e3:.1f} us on {nb} SMs; CTA resident {sum(busy.values()) / nb / span:.3f} of the span, "
          f"entry -> first score tile {sum(first.values()) / nb / span:.3f} of the span ({(rec[:, 2] - rec[:, 1]).double().mean() / 1e3:.2f} us per CTA); "
          f"gap between consecutive CTAs of an SM: mean {gaps.mean() / 1e3:.2f} us, median {gaps.median() / 1e3:.2f} us, max {gaps.max() / 1e3:.2f} us; "
          f"CTAs per SM {CTAS / nb:.1f}; mean CTA time {(rec[:, 3] - rec[:, 1]).double().mean() / 1e3:.2f} us; this CTA {int(own[3] - own[1]) / 1e3:.2f} us "
          f"(entry -> first tile {int(own[2] - own[1]) / 1e3:.2f} us), last CTA exit at {int(rec[:, 3].max() - g0) / 1e3:.1f} us, "
          f"earliest SM idle at {min(int(rec[rec[:, 0] == sm][:, 3].max()) for sm in busy) - g0:d} ns")
