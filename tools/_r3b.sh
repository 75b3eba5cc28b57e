mkdir -p gpurun_out/r3c
timeout 300 python tools/kernel_check.py decode > gpurun_out/r3c/kcheck_decode.log 2>&1; echo "kcheck exit $?"
grep -E "^\[FAIL\]|^==|Error|error" gpurun_out/r3c/kcheck_decode.log | tail -10
timeout 900 python -m pytest tests -m gpu -x -q -s -k "incremental or generate" > gpurun_out/r3c/pytest_decode.log 2>&1; echo "pytest exit $?"
grep -E "max=|passed|failed|Error|error|FAIL|assert" gpurun_out/r3c/pytest_decode.log | tail -30
timeout 600 python tools/bench_decode.py > gpurun_out/r3c/bench_decode.log 2>&1; tail -1 gpurun_out/r3c/bench_decode.log
timeout 600 python tools/bench_decode.py --prompt 1920 --new 128 > gpurun_out/r3c/bench_decode_long.log 2>&1; tail -1 gpurun_out/r3c/bench_decode_long.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:^(decode_|argmax_)" -s 800 -c 250 --csv --log-file gpurun_out/r3c/launches_decode.csv python tools/bench_decode.py --new 16 > gpurun_out/r3c/ncu.log 2>&1; echo "ncu exit $?"
