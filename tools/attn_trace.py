"""Timeline of one attention CTA (heaviest tile pair) from the traced build: clock64 deltas per KV block.
Usage: python tools/attn_trace.py [T]   (needs a B200; prints a table)"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kosmos-x_b200"))
from kosmosx import _abi, ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B, H = 8, 32
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").bfloat16()
q, k, v = qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:]
out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
buf = torch.zeros(4, 64, 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.attention(q, k, v, out, batch=B, heads=H, seq_len=T, causal=True, scale=0.125)
_abi.check(_abi.lib.kx_attn_set_trace(buf.data_ptr()), "kx_attn_set_trace")
ops.attention(q, k, v, out, batch=B, heads=H, seq_len=T, causal=True, scale=0.125)
torch.cuda.synchronize()
_abi.lib.kx_attn_set_trace(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
n = T // 128
print("softmax points: 0 iter start, 1 S ready, 2 S in regs (s_free), 3 max done, 4 exp done, 5 P stored, 6 p_full arrived")
print("MMA points: 0 iter start, 1 s_free seen, 2 next S issued, 3 v_full + p_full seen, 4 PV issued")
for role, name in ((0, "softmax A"), (1, "softmax B"), (2, "MMA A"), (3, "MMA B")):
    print(f"--- {name}")
    for j in range(n):
        row = t[role, j]
        if int(row.max()) == 0:
            continue
        rel = [(int(x) - t0) if int(x) > 0 else -1 for x in row]
        print(f"blk {j:2d}: " + " ".join(f"{x:7d}" for x in rel))
