"""Timeline of the staged GEMM epilogue (warp 4 of CTA 0) from the -DKX_GEMM_TRACE build.
Build:  KX_BUILD_TRACE=1 bash kosmos-x_b200/build.sh
Run:    KX_LIB=kosmos-x_b200/lib/libkosmosx_sm100_trace.so python tools/gemm_trace.py <fc1|out|qkv|fc2>"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kosmos-x_b200"))
from kosmosx import _abi, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
M = 16384
shape = {"fc1": (8192, 2048), "out": (2048, 2048), "qkv": (6144, 2048), "fc2": (2048, 8192)}[which]
N, K = shape
dev = "cuda"
a = torch.randn(M, K, device=dev).bfloat16()
w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
bias = torch.randn(N, device=dev)
c = w.float().sum(1)
tiles_in = {"fc1": 8, "out": 32, "qkv": 8, "fc2": 32}[which]
part = torch.randn(tiles_in, M, 2, device=dev).abs() + 1.0
part[:, :, 1] = part[:, :, 0] ** 2 * 4 + 10
st = torch.zeros((N + 127) // 128, M, 2, device=dev)
kw = dict(bias=bias, ln=(part, c, K, 1e-5))
if which == "fc1":
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); kw.update(act=_abi.KX_ACT_GELU, stats_out=st)
elif which == "qkv":
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    tabs = torch.randn(4, 2048, 32, device=dev)
    kw.update(xpos=(tabs[0], tabs[1], tabs[2], tabs[3]), seq_len=2048)
else:
    out = torch.randn(M, N, device=dev); xb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    kw.update(res=out, stats_out=st, out2=xb)
for _ in range(3):
    ops.gemm(a, w, out, **kw)
buf = torch.zeros(16, 8, 8, dtype=torch.int64, device=dev)
lib = _abi.lib
lib.kx_gemm_set_trace.restype = C.c_int
lib.kx_gemm_set_trace.argtypes = [C.c_void_p]
assert lib.kx_gemm_set_trace(buf.data_ptr()) == 0
ops.gemm(a, w, out, **kw)
torch.cuda.synchronize()
lib.kx_gemm_set_trace(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
print(f"{which}: M={M} N={N} K={K}; points per chunk: 3 chunk start, 4 after store-wait/res-issue, 5 TMEM data ready, 6 math done,"
      " 7 residual landed; chunk0 extra: 0 tile start (overwritten at tile end), 1 vectors staged, 2 accumulator ready")
for tile in range(8):
    if int(t[tile].max()) == 0:
        continue
    r0 = t[tile, 0]
    print(f"tile {tile}: end {int(r0[0]) - t0:7d} vec {int(r0[1]) - t0:7d} acc-ready {int(r0[2]) - t0:7d}")
    for ch in range(8):
        r = t[tile, ch]
        print("    chunk %d: " % ch + " ".join(f"{(int(x) - t0) if int(x) > 0 else -1:7d}" for x in r[3:8]))
