set -x
mkdir -p gpurun_out
MP_RTC_TS=1 timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | grep "rtc ts" | sed -n 2,12p
timeout 200 python scripts/rtc_time.py 2>&1 | grep -E "lstm_rec|gemm_f16|split|linear"
MP_NO_FUSED_SPLIT=1 timeout 200 python scripts/rtc_time.py 2>&1 | grep -E "lstm_rec|gemm_f16|split|linear"
timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | tail -1
bash scripts/gpu_sanitize.sh
