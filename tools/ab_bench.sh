#!/usr/bin/env bash
# A/B on ONE box: bench.py with two builds of the library, interleaved.  Usage: tools/ab_bench.sh <tag> <libA> <libB> [rounds]
cd "$(dirname "$0")/.."
tag="$1"; A="$2"; B="$3"; rounds="${4:-2}"
out="gpurun_out/$tag"; mkdir -p "$out"
for r in $(seq 1 "$rounds"); do
  for v in A B; do
    lib="${!v}"
    KX_LIB="$lib" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --train-leg 0 > "$out/bench_${v}_$r.json" 2> "$out/bench_${v}_$r.err" || tail -3 "$out/bench_${v}_$r.err"
    python - "$out/bench_${v}_$r.json" "$v" "$lib" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
g = d["gemm_launch_types"]
print(sys.argv[2], sys.argv[3].split("/")[-1], "ms/step %.2f" % d["ms_per_step"], "tok/s %.0f" % d["value"], "clk", d["clocks"]["sm_mhz"],
      "dec_block %.3f" % d["decoder_block"]["ms"], " | ".join("%s %.0fus" % (k.split("+")[0][6:], v["us_per_launch"]) for k, v in list(g.items())[:5]),
      "attn %.2f" % d["breakdown"]["attn_causal"]["ms_per_step"])
PY
  done
done | tee "$out/ab.txt"
