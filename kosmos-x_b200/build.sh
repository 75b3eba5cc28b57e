#!/usr/bin/env bash
# Build libkosmosx_sm100.so in-tree (sm_100a only).  Usage: kosmos-x_b200/build.sh [-v]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/lib"
mkdir -p "$out" "$here/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -Wno-deprecated-gpu-targets -Xptxas -v)
objs=()
pids=()
for src in runtime gemm accurate preprocess attention attention_pp attention_bwd elementwise train perceiver_attn decode decode_step; do
  obj="$here/build/$src.o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$here/csrc/$src.cu" -nt "$obj" || "$here/csrc/ptx.cuh" -nt "$obj" || "$here/csrc/kx_internal.h" -nt "$obj" || "$here/../include/kosmosx_b200.h" -nt "$obj" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$here/csrc/$src.cu" -o "$obj" > "$here/build/$src.log" 2>&1 || { cat "$here/build/$src.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$out/libkosmosx_sm100.so" "${objs[@]}" -cudart static
if [[ "${KX_BUILD_TRACE:-0}" == "1" ]]; then      # profiling variant: GEMM epilogue clock64 stamps (tools/gemm_trace.py)
  "$NVCC" "${FLAGS[@]}" -DKX_GEMM_TRACE -c "$here/csrc/gemm.cu" -o "$here/build/gemm_trace.o" > "$here/build/gemm_trace.log" 2>&1 || { cat "$here/build/gemm_trace.log"; exit 1; }
  tobjs=("${objs[@]/$here\/build\/gemm.o/$here/build/gemm_trace.o}")
  "$NVCC" -shared -o "$out/libkosmosx_sm100_trace.so" "${tobjs[@]}" -cudart static
fi
if [[ "${1:-}" == "-v" ]]; then grep -h -E "registers|spill|error|warning" "$here"/build/*.log || true; fi
echo "built $out/libkosmosx_sm100.so"
