#!/usr/bin/env bash
# 2-GPU validation of bench.py under torchrun (own arm + reference arm).  Usage: tools/run_2gpu.sh <tag> [N]
cd "$(dirname "$0")/.."
tag="${1:-r1k}"; n="${2:-2}"
out="gpurun_out/$tag"; mkdir -p "$out"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > "$out/bench_${n}gpu.json" 2> "$out/bench_${n}gpu.err"
echo "bench exit $?"; tail -3 "$out/bench_${n}gpu.err"; cut -c1-700 "$out/bench_${n}gpu.json"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $n --steps 1 --warmup 0 --impl reference > "$out/ref_${n}gpu.json" 2> "$out/ref_${n}gpu.err"
echo "ref exit $?"; tail -3 "$out/ref_${n}gpu.err"; cut -c1-500 "$out/ref_${n}gpu.json"
