#!/usr/bin/env bash
# Round-2 evidence in ONE gpurun call (single GPU): per-kernel ncu tables of the forward and of the training step, and
# `ncu --set full` pages of the final kernels (raw pages exported to CSV on the box; the .ncu-rep files stay there).
# Usage: tools/gpu_evidence.sh <tag> [tables] [full_fwd] [full_train]
cd "$(dirname "$0")/.."
tag="${1:-r2ev}"; shift || true
what="${*:-tables full_fwd full_train}"
out="gpurun_out/$tag"; mkdir -p "$out"
M6="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
full() {   # full <name> <kernel regex> <skip> <count> <command...>
  local name="$1" rx="$2" skip="$3" cnt="$4"; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s "$skip" -c "$cnt" -o "/tmp/$name" -f "$@" > "$out/ncu_$name.log" 2>&1
  echo "ncu $name exit $?" | tee -a "$out/summary.txt"
  ncu -i "/tmp/$name.ncu-rep" --page raw --csv > "$out/r2_ncu_${name}_raw.csv" 2>> "$out/ncu_$name.log"
}
for w in $what; do
  case $w in
    tables)
      timeout 900 ncu --metrics "$M6" --clock-control none --csv --log-file "$out/kernel_table_train.csv" python tools/profile_step.py --train --steps 2 --layers 2 --vit-layers 2 > "$out/kernel_table_train.log" 2>&1; echo "table train exit $?" | tee -a "$out/summary.txt"
      timeout 600 ncu --metrics "$M6" --clock-control none --csv --log-file "$out/kernel_table_fwd.csv" python tools/profile_step.py --steps 2 --layers 2 --vit-layers 2 > "$out/kernel_table_fwd.log" 2>&1; echo "table fwd exit $?" | tee -a "$out/summary.txt";;
    full_fwd)
      # forward, 2 decoder layers + 1 ViT layer: 16 GEMM launches precede decoder layer 0 (patch, 4 ViT, 10 resampler, image_proj)
      full gemm_decoder_layer gemm_bf16 16 4 python tools/profile_step.py --steps 1 --layers 2 --vit-layers 1
      full attn_fwd_causal "attn_pp_kernel<\(bool\)1" 0 1 python tools/profile_step.py --steps 1 --layers 2 --vit-layers 1
      full gemm_lm_head gemm_bf16 24 1 python tools/profile_step.py --steps 1 --layers 2 --vit-layers 1;;
    full_train)
      full attn_bwd attn_bwd_kernel 1 1 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1
      full attn_fwd_dropout "attn_pp_kernel<\(bool\)1, \(bool\)1, \(bool\)0, \(bool\)1" 0 1 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1
      full ln_bwd_gelu "layernorm_bwd_kernel<__nv_bfloat16, \(bool\)0, \(int\)1024" 0 1 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1
      full gemm_dgrad_wgrad gemm_bf16 34 4 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1
      full attn_dropout_masks attn_dropout_mask_kernel 0 1 python tools/profile_step.py --train --steps 1 --layers 2 --vit-layers 1;;
  esac
done
ls -la "$out" | tail -20
