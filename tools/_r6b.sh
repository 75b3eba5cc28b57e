mkdir -p gpurun_out/r6f
KX_STEP_TRACE=1 timeout 300 python tools/bench_decode.py --mode one > gpurun_out/r6f/bench_one.log 2>&1; grep "^trace" gpurun_out/r6f/bench_one.log; tail -1 gpurun_out/r6f/bench_one.log | cut -c1-200; tail -1 gpurun_out/r6f/bench_one.log | grep -o '"path[^,]*'
timeout 300 python tools/bench_decode.py --mode one --prompt 1920 > gpurun_out/r6f/bench_one_long.log 2>&1; tail -1 gpurun_out/r6f/bench_one_long.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -x -q -k "incremental or generate" > gpurun_out/r6f/pytest.log 2>&1; echo "pytest exit $?"; tail -1 gpurun_out/r6f/pytest.log
