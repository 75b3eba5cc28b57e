cd /root/repo 2>/dev/null || true
out=gpurun_out/r2exp; mkdir -p $out
run() { name=$1; shift; env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 2 --workload train --steps 5 --warmup 3 --dropout 0 --reduce-bf16 1 > $out/$name.json 2> $out/$name.err; python - $out/$name.json $name <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); b=d["breakdown"]
print(sys.argv[2], "ms/step %.1f" % d["ms_per_step"], {k: round(b[k]["ms"],1) for k in ("gemm fwd","gemm wgrad","gemm dgrad","layernorm_bwd","attn_bwd")})
PY
}
run base KX_X=1
run nooverlap KX_BENCH_OVERLAP=0
run cap4_ctas140 NCCL_MAX_CTAS=4 KX_BENCH_BWD_CTAS=140
run cap4_ctas132 NCCL_MAX_CTAS=4 KX_BENCH_BWD_CTAS=132
run cap4_only NCCL_MAX_CTAS=4
run cap8_ctas132 NCCL_MAX_CTAS=8 KX_BENCH_BWD_CTAS=132
