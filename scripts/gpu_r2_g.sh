set -x
mkdir -p gpurun_out
MP_RTC_TS=1 timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | grep "rtc ts" | sed -n 2,9p
timeout 100 python scripts/rtc_debug.py 70 24 f16 2>&1 | tail -1
timeout 200 python scripts/rtc_time.py 2>&1 | grep -E "lstm_rec|gemm_f16|split|linear"
timeout 900 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-260 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_g.json; tail -5 gpurun_out/bench_g.err
./scripts/sanitizer/dsmem_bulk_repro
timeout 300 compute-sanitizer --tool memcheck ./scripts/sanitizer/dsmem_bulk_repro > gpurun_out/san_repro_memcheck.log 2>&1; echo "repro memcheck exit $?"
grep -E "^========= [A-Z]|ERROR SUMMARY|dsmem bulk|not located" gpurun_out/san_repro_memcheck.log | sort | uniq -c | sort -rn | head -8
