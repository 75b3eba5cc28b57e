#!/usr/bin/env bash
# Final single-GPU evidence of the second session of round 2 (GPU-minute budget: two short calls).
# Usage: tools/final_round_s2.sh <tag> bench|ncu
cd "$(dirname "$0")/.."
tag="${1:-r2s2}"; what="${2:-bench}"; out="gpurun_out/$tag"; mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$out/gpu.txt" 2>&1; nproc >> "$out/gpu.txt"
M6="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
full() {   # full <name> <kernel regex> <skip> <count> <command...>
  local name="$1" rx="$2" skip="$3" cnt="$4"; shift 4
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s "$skip" -c "$cnt" -o "/tmp/$name" -f "$@" > "$out/ncu_$name.log" 2>&1
  echo "ncu $name exit $?" | tee -a "$out/summary.txt"
  ncu -i "/tmp/$name.ncu-rep" --page raw --csv > "$out/r2s2_ncu_${name}_raw.csv" 2>> "$out/ncu_$name.log"
}
if [[ "$what" == "bench" ]]; then
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke exit $?" | tee -a "$out/summary.txt"; tail -1 "$out/smoke.log"
  timeout 600 python bench.py --steps 10 --warmup 3 > "$out/bench.json" 2> "$out/bench.err"; echo "bench exit $?" | tee -a "$out/summary.txt"
  timeout 300 python bench.py --workload train --steps 5 --warmup 3 > "$out/bench_train.json" 2> "$out/bench_train.err"; echo "bench train (dropout 0.1) exit $?" | tee -a "$out/summary.txt"
  timeout 300 python bench.py --workload train --steps 5 --warmup 3 --dropout 0 > "$out/bench_train_nodropout.json" 2> "$out/bench_train_nodropout.err"; echo "bench train (no dropout) exit $?" | tee -a "$out/summary.txt"
  for c in train_elementwise attn_bwd; do
    timeout 300 compute-sanitizer --tool memcheck python tools/kernel_check.py $c > "$out/sanitizer_memcheck_$c.log" 2>&1; echo "memcheck $c: $(grep 'ERROR SUMMARY' "$out/sanitizer_memcheck_$c.log")" | tee -a "$out/summary.txt"
  done
else
  timeout 600 ncu --metrics "$M6" --clock-control none --csv --log-file "$out/kernel_table_train.csv" python tools/profile_step.py --train --steps 2 --layers 2 --vit-layers 2 > "$out/kernel_table_train.log" 2>&1; echo "table train exit $?" | tee -a "$out/summary.txt"
  full ln_gelu_bwd_wide ln_gelu_bwd_wide_kernel 0 1 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1
  full attn_bwd attn_bwd_kernel 1 1 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1
  full act_layernorm_fwd act_layernorm_fwd_kernel 0 1 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1
fi
cat "$out/summary.txt"
