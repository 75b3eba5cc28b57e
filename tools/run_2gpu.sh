#!/usr/bin/env bash
# N-GPU validation of bench.py under torchrun (own arm incl. the train_step leg, the training workload, reference arm).
# Usage: tools/run_2gpu.sh <tag> [N]
cd "$(dirname "$0")/.."
tag="${1:-r1k}"; n="${2:-2}"
out="gpurun_out/$tag"; mkdir -p "$out"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > "$out/bench_${n}gpu.json" 2> "$out/bench_${n}gpu.err"
echo "bench exit $?"; tail -3 "$out/bench_${n}gpu.err"; python - "$out/bench_${n}gpu.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("forward: n_gpus", d["n_gpus"], "tok/s %.0f" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
t = d.get("train_step", {})
print("train_step:", {k: t.get(k) for k in ("value", "ms_per_step", "n_gpus", "grad_all_reduce", "loss_first", "loss_last", "error")})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 5 --warmup 3 --workload train > "$out/train_${n}gpu.json" 2> "$out/train_${n}gpu.err"
echo "train exit $?"; tail -3 "$out/train_${n}gpu.err"; cut -c1-900 "$out/train_${n}gpu.json"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $n --steps 1 --warmup 0 --impl reference > "$out/ref_${n}gpu.json" 2> "$out/ref_${n}gpu.err"
echo "ref exit $?"; tail -3 "$out/ref_${n}gpu.err"; cut -c1-400 "$out/ref_${n}gpu.json"
