# round 2, visit D: housekeeping moved to the copy warp + exchange stamps; persistent GEMM with TMA-store epilogue
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -q -rP -k "tensor_core_gemm or fp16_split" > gpurun_out/pytest_gemm.log 2>&1; echo "gemm exit $?"
grep -E "f16x3|passed|failed" gpurun_out/pytest_gemm.log | cut -c1-200 | tail -12
timeout 300 python scripts/time_gemm16.py > gpurun_out/time_gemm16.log 2>&1; grep "mode=3 gemm" gpurun_out/time_gemm16.log
for v in "" "MP_RF16_ACT=exact"; do
  echo "== variant: $v"
  env $v timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | tail -1
  env $v timeout 100 python scripts/rtc_debug.py 70 24 f16 2>&1 | tail -1
  env $v timeout 200 python scripts/rtc_time.py 2>&1 | grep -E "lstm_rec|gemm_f16"
  env $v MP_RTC_TS=1 timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | grep "rtc ts" | sed -n 2,5p
done
timeout 900 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-260 | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --min-seconds 1 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_d.json; tail -5 gpurun_out/bench_d.err
