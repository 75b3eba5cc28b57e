#!/usr/bin/env bash
# Multi-GPU measurement round (one gpurun --gpus N call): the bench line of every workload under torchrun + the NCCL correctness check.
# Usage: tools/multi_gpu_round.sh <N> <tag>
cd "$(dirname "$0")/.."
N="${1:-2}"; tag="${2:-r2mg}"
out="gpurun_out/$tag"; mkdir -p "$out"
run() {   # run <name> <bench args...>
  local name="$1"; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus "$N" "$@" > "$out/bench_${name}_${N}gpu.json" 2> "$out/bench_${name}_${N}gpu.err"
  echo "bench $name N=$N exit $?" | tee -a "$out/summary.txt"
  python - "$out/bench_${name}_${N}gpu.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e", {})
    print("   value %.0f tok/s, %.2f ms/step; e2e %.0f tok/s (%.2f ms), d2h %.1f GB/s per GPU; train %s" % (
        d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step", 0), e.get("d2h_gbs_per_gpu_all_ranks_copying", 0),
        {k: d.get("train_step", {}).get(k) for k in ("ms_per_step", "value")} if "train_step" in d else "-"))
except Exception as ex:
    print("   (no line:", ex, ")")
PY
}
run c3 --steps 10 --warmup 3
run strict --workload strict --steps 10 --warmup 3 --train-leg 0
run c5 --workload c5 --steps 10 --warmup 3 --train-leg 0
run train_fp32reduce_nodropout --workload train --steps 5 --warmup 3 --dropout 0 --reduce-bf16 0
run train_bf16reduce_nodropout --workload train --steps 5 --warmup 3 --dropout 0 --reduce-bf16 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29555 tools/dp_check.py > "$out/dp_check_${N}gpu.log" 2>&1
echo "dp_check N=$N exit $?" | tee -a "$out/summary.txt"; grep -E "overlap=|bridge|DP_CHECK" "$out/dp_check_${N}gpu.log"
