mkdir -p gpurun_out/r4a
( time timeout 1200 python bench.py > gpurun_out/r4a/bench.json 2> gpurun_out/r4a/bench.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4a/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d.get('clocks'))
print('train', {k:d['train_step'].get(k) for k in ('value','ms_per_step','error')})
print('decode', d.get('decode'))
print('cpu', d.get('cpu_baseline'))
PY
tail -3 gpurun_out/r4a/bench.err
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4a/bench_ref.json 2> gpurun_out/r4a/bench_ref.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/r4a/bench_ref.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
