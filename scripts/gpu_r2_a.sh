# round 2, visit A: whole gpu suite with the parity prints, smoke, the recurrence's per-step stamps (baseline for the f16 work)
set -x
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
# the new fp16-split projection first, alone and under a short timeout; if it is not green the rest runs on the TF32 kernel
timeout 120 python -m pytest tests/test_gpu_parity.py -q -rP -k "tensor_core_gemm or fp16_split" > gpurun_out/pytest_gemm.log 2>&1; g=$?; echo "gemm exit $g"
grep -E "^gemm M|passed|failed|rror" gpurun_out/pytest_gemm.log | cut -c1-200 | tail -30
if [ $g -ne 0 ]; then export MP_GEMM=tf32; echo "FALLING BACK TO MP_GEMM=tf32"; fi
# the new fp16-split recurrence alone (70 sequences x 24 steps and 256 x 40) against the FFMA kernel, short timeout
timeout 120 python scripts/rtc_debug.py 70 24 f16 > gpurun_out/rec_f16_check.log 2>&1; r=$?; tail -2 gpurun_out/rec_f16_check.log
timeout 120 python scripts/rtc_debug.py 256 40 f16 >> gpurun_out/rec_f16_check.log 2>&1; r2=$?; tail -1 gpurun_out/rec_f16_check.log
ok=$(python - <<'PY'
import re
errs = [float(m.group(1)) for m in re.finditer(r"max \|tc - ffma\| ([0-9.e+-]+)", open('gpurun_out/rec_f16_check.log').read())]
print(1 if len(errs) == 2 and max(errs) < 2e-6 else 0)
PY
)
if [ $r -ne 0 ] || [ $r2 -ne 0 ] || [ "$ok" != "1" ]; then export MP_REC_IMPL=tf32; echo "FALLING BACK TO MP_REC_IMPL=tf32"; fi
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "^\[(angle|flat|tran)\]|passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-260 | tail -70
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
bash scripts/gpu_rtc_ts.sh
timeout 300 python scripts/time_gemm.py > gpurun_out/time_gemm.log 2>&1; cat gpurun_out/time_gemm.log
timeout 300 python scripts/time_gemm16.py > gpurun_out/time_gemm16.log 2>&1; cat gpurun_out/time_gemm16.log
