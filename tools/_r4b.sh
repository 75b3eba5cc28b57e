mkdir -p gpurun_out/r4b
timeout 300 python tools/kernel_check.py decode > gpurun_out/r4b/kcheck.log 2>&1; echo "kcheck exit $?"; grep -E "^\[FAIL\]|^==" gpurun_out/r4b/kcheck.log
timeout 600 python -m pytest tests -m gpu -x -q -s -k "incremental or generate" > gpurun_out/r4b/pytest.log 2>&1; echo "pytest exit $?"
grep -E "max=|passed|failed|Error|error|FAIL|assert" gpurun_out/r4b/pytest.log | tail -30
timeout 300 python tools/bench_decode.py --mode graph > gpurun_out/r4b/bench_graph.log 2>&1; tail -1 gpurun_out/r4b/bench_graph.log | cut -c1-330
timeout 300 python tools/bench_decode.py --mode graph --prompt 1920 > gpurun_out/r4b/bench_graph_long.log 2>&1; tail -1 gpurun_out/r4b/bench_graph_long.log | cut -c1-330
