"""Timeline of one CTA of the attention-backward kernel (kx_attn_bwd_set_trace): clock64 stamps of compute thread 0
and of the MMA-issuing thread for CTA 0 (batch 0, head 0, key block 0: 16 query blocks at T = 2048).
    python tools/attn_bwd_trace.py > gpurun_out/attn_bwd_trace.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kosmos-x_b200"))
from kosmosx import _abi, ops  # noqa: E402

dev = torch.device("cuda")
B, H, T = 8, 32, 2048
D, M = H * 64, B * T
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
buf = torch.full((2, 32, 16), -1, dtype=torch.int64, device=dev)
_abi.check(_abi.lib.kx_attn_bwd_set_trace(buf.data_ptr()), "kx_attn_bwd_set_trace")
run()
torch.cuda.synchronize()
_abi.lib.kx_attn_bwd_set_trace(None)
t = buf.cpu()
t0 = int(t[t >= 0].min())
print("compute thread 0 points: 0 iter start, 1 S^T ready, 2 dP^T ready, 4 P^T / dS^T computed, 5 previous dQ MMA done (smem dS^T, P^T free),"
      " 6 P^T / dS^T stored + arrived")
print("MMA thread points: 0 loop top, 1 next S^T + dP^T issued, 2 P^T / dS^T seen, 3 dV + dK issued, 4 previous dQ read out, 5 dQ issued")
for role, name in ((0, "compute"), (1, "MMA")):
    print("---", name)
    for it in range(16):
        row = t[role, it]
        print(f"it {it:2d}: " + " ".join(f"{int(v) - t0:7d}" if v >= 0 else "     -1" for v in row[:7 if role == 0 else 6]))
