set -x
mkdir -p gpurun_out
MP_RTC_TS=1 timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | grep "rtc ts" | sed -n 2,9p
timeout 200 python scripts/rtc_time.py 2>&1 | grep -E "lstm_rec|gemm_f16|split|linear"
MP_REC_IMPL=tf32 timeout 200 python scripts/rtc_time.py 2>&1 | grep -E "lstm_rec"
timeout 300 python scripts/time_gemm16.py > gpurun_out/time_gemm16.log 2>&1; grep "mode=3 gemm" gpurun_out/time_gemm16.log
timeout 900 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-260 | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --min-seconds 1 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_f.json; tail -5 gpurun_out/bench_f.err
./scripts/sanitizer/dsmem_bulk_repro
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool ./scripts/sanitizer/dsmem_bulk_repro > gpurun_out/san_repro_$tool.log 2>&1; echo "repro $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|dsmem bulk|not located" gpurun_out/san_repro_$tool.log | sort | uniq -c | sort -rn | head -8
done
MP_REC_IMPL=f16 timeout 600 compute-sanitizer --tool memcheck --print-limit 2000 python scripts/rtc_debug.py 20 5 f16 > gpurun_out/san_rec_f16_memcheck.log 2>&1
grep -o "Device Frame: void mp::<unnamed>::[a-z_0-9]*" gpurun_out/san_rec_f16_memcheck.log | sort | uniq -c
