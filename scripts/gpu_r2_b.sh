# round 2, visit B: full suite on the fp16-split kernels, A/B of the recurrence epilogue variants, ncu captures, first run of the new bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-260 | tail -12
for v in "" "MP_RF16_ACT=fast" "MP_RF16_RAGGED=1" "MP_RF16_ACT=fast MP_RF16_RAGGED=1" "MP_REC_IMPL=tf32"; do
  echo "== variant: $v"; env $v timeout 200 python scripts/rtc_time.py 2>&1 | tail -4
  env $v timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | tail -1
done
timeout 600 python bench.py --steps 5 --warmup 3 --min-seconds 1 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_f16 -c 1 -o gpurun_out/prof_rec_f16_b256 python scripts/prof_one.py --batch 256 --passes 1 --tile 64 > gpurun_out/prof_a.log 2>&1; echo "ncu rec exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 2 -c 1 -o gpurun_out/prof_gemm_f16 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_b.log 2>&1; echo "ncu gemm exit $?"
