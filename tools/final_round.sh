#!/usr/bin/env bash
# Final single-GPU round of a build: GPU tests, the default bench line, the training line without dropout, compute-sanitizer over the
# kernels added in round 2, then the ncu evidence (tools/gpu_evidence.sh).  Usage: tools/final_round.sh <tag>
cd "$(dirname "$0")/.."
tag="${1:-r2final}"; out="gpurun_out/$tag"; mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$out/gpu.txt" 2>&1; nproc >> "$out/gpu.txt"
timeout 900 python -m pytest tests -m gpu -q -s > "$out/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$out/summary.txt"; tail -1 "$out/pytest_gpu.log"
timeout 900 python bench.py --steps 10 --warmup 3 > "$out/bench.json" 2> "$out/bench.err"; echo "bench exit $?" | tee -a "$out/summary.txt"
timeout 400 python bench.py --workload train --steps 5 --warmup 3 --dropout 0 > "$out/bench_train_nodropout.json" 2> "$out/bench_train_nodropout.err"; echo "bench train exit $?" | tee -a "$out/summary.txt"
timeout 400 python bench.py --workload train --steps 5 --warmup 3 > "$out/bench_train.json" 2> "$out/bench_train.err"; echo "bench train (dropout 0.1) exit $?" | tee -a "$out/summary.txt"
timeout 400 python bench.py --workload train --steps 5 --warmup 3 --clip-last 1 > "$out/bench_train_cliplast.json" 2> "$out/bench_train_cliplast.err"; echo "bench train clip-last exit $?" | tee -a "$out/summary.txt"
timeout 400 python bench.py --workload train --steps 5 --warmup 3 --recompute 1 > "$out/bench_train_recompute.json" 2> "$out/bench_train_recompute.err"; echo "bench train recompute exit $?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --workload strict --steps 10 --warmup 3 --train-leg 0 --decode-leg 0 --no-cpu > "$out/bench_strict.json" 2> "$out/bench_strict.err"; echo "bench strict exit $?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --train-leg 0 --decode-leg 0 --no-cpu > "$out/bench_c5.json" 2> "$out/bench_c5.err"; echo "bench c5 exit $?" | tee -a "$out/summary.txt"
for c in accurate attn_dropout perceiver_attn; do
  timeout 600 compute-sanitizer --tool memcheck python tools/kernel_check.py $c > "$out/sanitizer_memcheck_$c.log" 2>&1; echo "memcheck $c: $(grep 'ERROR SUMMARY' "$out/sanitizer_memcheck_$c.log")" | tee -a "$out/summary.txt"
done
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "resize or raw_pictures" > "$out/sanitizer_memcheck_resize.log" 2>&1; echo "memcheck resize: $(grep 'ERROR SUMMARY' "$out/sanitizer_memcheck_resize.log")" | tee -a "$out/summary.txt"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_train.py -m gpu -q -k "clip_last_layer_fine_tuning_gradients" > "$out/sanitizer_memcheck_clip_last.log" 2>&1; echo "memcheck clip-last: $(grep 'ERROR SUMMARY' "$out/sanitizer_memcheck_clip_last.log")" | tee -a "$out/summary.txt"
timeout 600 compute-sanitizer --tool racecheck python tools/kernel_check.py accurate > "$out/sanitizer_racecheck_accurate.log" 2>&1; echo "racecheck accurate: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' "$out/sanitizer_racecheck_accurate.log" | tail -1)" | tee -a "$out/summary.txt"
bash tools/gpu_evidence.sh "$tag" tables full_fwd full_train > "$out/evidence.log" 2>&1
cat "$out/summary.txt"
