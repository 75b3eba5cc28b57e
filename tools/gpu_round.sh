#!/usr/bin/env bash
# One gpurun call: GPU parity tests, bench, ncu launch list + full captures.  Logs under gpurun_out/<tag>/.
# Usage: tools/gpu_round.sh <tag> [tests] [bench] [launches] [ncu_gemm] [ncu_attn] [smoke] [decode] ...
cd "$(dirname "$0")/.."
tag="${1:-r1}"; shift || true
what="${*:-smoke tests bench launches ncu_gemm ncu_attn}"
out="gpurun_out/$tag"; mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$out/gpu.txt" 2>&1
nproc >> "$out/gpu.txt"; free -g | head -2 >> "$out/gpu.txt"
for w in $what; do
  case $w in
    smoke)    timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke exit $?" | tee -a "$out/summary.txt"; tail -2 "$out/smoke.log";;
    tests)    timeout 1500 python -m pytest tests -m gpu -x -q -s > "$out/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$out/summary.txt"; grep -E "max=|passed|failed|Error|error|FAIL|agreement" "$out/pytest_gpu.log" | tail -40;;
    bench)    timeout 900 python bench.py --steps 10 --warmup 3 > "$out/bench.json" 2> "$out/bench.err"; echo "bench exit $?" | tee -a "$out/summary.txt"; cat "$out/bench.json"; tail -5 "$out/bench.err";;
    benchng)  timeout 900 python bench.py --steps 10 --warmup 3 --graph 0 --no-cpu > "$out/bench_nograph.json" 2> "$out/bench_nograph.err"; echo "bench(nograph) exit $?" | tee -a "$out/summary.txt"; cat "$out/bench_nograph.json"; tail -5 "$out/bench_nograph.err";;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(attn_|broadcast_rows|embed_splice|gemm_bf16|im2col|layernorm|perceiver_x)" -s 413 -c 413 --csv --log-file "$out/launches.csv" python tools/profile_step.py --steps 3 > "$out/launches.log" 2>&1; echo "launches exit $?" | tee -a "$out/summary.txt"; tail -3 "$out/launches.log";;
    ncu_gemm) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 16 -c 4 -o "$out/prof_gemm" -f python tools/profile_step.py --steps 1 --layers 2 --vit-layers 1 > "$out/ncu_gemm.log" 2>&1; echo "ncu_gemm exit $?" | tee -a "$out/summary.txt"; tail -3 "$out/ncu_gemm.log";;
    ncu_attn) timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 1 -c 1 -o "$out/prof_attn" -f python tools/profile_step.py --steps 1 --layers 2 --vit-layers 1 > "$out/ncu_attn.log" 2>&1; echo "ncu_attn exit $?" | tee -a "$out/summary.txt"; tail -3 "$out/ncu_attn.log";;
    kcheck)   bash tools/run_kernel_checks.sh bench;;
    preprocess) for c in embed preprocess bench_preprocess; do timeout 100 python tools/kernel_check.py $c > "$out/kcheck_$c.log" 2>&1; echo "kcheck $c exit $?" | tee -a "$out/summary.txt"; done; grep "^preprocess" "$out/kcheck_bench_preprocess.log";
                timeout 100 compute-sanitizer --tool memcheck python tools/kernel_check.py preprocess > "$out/sanitizer_preprocess.log" 2>&1; grep "ERROR SUMMARY" "$out/sanitizer_preprocess.log";;
    bench_vit) timeout 150 python tools/kernel_check.py bench_vit > "$out/bench_vit.log" 2>&1; grep "^gemm" "$out/bench_vit.log";;
    train_tests) timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q -s > "$out/pytest_train.log" 2>&1; echo "train tests exit $?" | tee -a "$out/summary.txt"; grep -E "bridge|accumulated|passed|failed" "$out/pytest_train.log" | tail -8;;
    decode)   timeout 300 python tools/kernel_check.py decode > "$out/kcheck_decode.log" 2>&1; echo "kcheck decode exit $?" | tee -a "$out/summary.txt";
              timeout 600 python -m pytest tests -m gpu -x -q -s -k "incremental or generate" > "$out/pytest_decode.log" 2>&1; echo "pytest decode exit $?" | tee -a "$out/summary.txt"; grep -E "max=|passed|failed" "$out/pytest_decode.log" | tail -20;
              for mode in one graph; do KX_STEP_TRACE=1 timeout 300 python tools/bench_decode.py --mode $mode > "$out/bench_decode_$mode.log" 2>&1; grep "^trace" "$out/bench_decode_$mode.log"; tail -1 "$out/bench_decode_$mode.log" | cut -c1-400; done;;
    kernel_table) timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none --csv --log-file "$out/kernel_table.csv" python tools/profile_step.py --train --steps 2 --layers 2 --vit-layers 2 > "$out/kernel_table.log" 2>&1; echo "kernel_table exit $?" | tee -a "$out/summary.txt"; tail -2 "$out/kernel_table.log";
                  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none --csv --log-file "$out/kernel_table_fwd.csv" python tools/profile_step.py --steps 2 --layers 2 --vit-layers 2 > "$out/kernel_table_fwd.log" 2>&1; echo "kernel_table_fwd exit $?" | tee -a "$out/summary.txt";;
    bench_train) timeout 900 python bench.py --workload train --steps 5 --warmup 3 > "$out/bench_train.json" 2> "$out/bench_train.err"; echo "bench_train exit $?" | tee -a "$out/summary.txt"; cut -c1-400 "$out/bench_train.json"; tail -3 "$out/bench_train.err";;
    launches_train) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file "$out/launches_train.csv" python tools/profile_step.py --train --steps 3 > "$out/launches_train.log" 2>&1; echo "launches_train exit $?" | tee -a "$out/summary.txt"; tail -3 "$out/launches_train.log";;
    ncu_attn_bwd) timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 2 -c 1 -o "$out/prof_attn_bwd" -f python tools/kernel_check.py bench_train > "$out/ncu_attn_bwd.log" 2>&1; echo "ncu_attn_bwd exit $?" | tee -a "$out/summary.txt";;
    ncu_wgrad) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 16 -c 1 -o "$out/prof_wgrad" -f python tools/kernel_check.py bench_train > "$out/ncu_wgrad.log" 2>&1; echo "ncu_wgrad exit $?" | tee -a "$out/summary.txt";
               timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 1 -o "$out/prof_dgrad" -f python tools/kernel_check.py bench_train > "$out/ncu_dgrad.log" 2>&1; echo "ncu_dgrad exit $?" | tee -a "$out/summary.txt";;
  esac
done
