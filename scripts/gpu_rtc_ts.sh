# per-step clock64 stamps of the tcgen05 recurrence (block (0,0)) + isolated launch time of one bidirectional layer
set -x
mkdir -p gpurun_out
MP_RTC_TS=1 timeout 300 python scripts/rtc_debug.py 256 40 > gpurun_out/rtc_ts.log 2>&1; grep -E "rtc ts|max" gpurun_out/rtc_ts.log | head -14
MP_RTC_TS=1 timeout 300 python scripts/rtc_debug.py 256 40 tf32 > gpurun_out/rtc_ts_tf32.log 2>&1; grep -E "rtc ts|max" gpurun_out/rtc_ts_tf32.log | head -14
timeout 300 python scripts/rtc_time.py > gpurun_out/rtc_time.log 2>&1
cat gpurun_out/rtc_time.log | tail -8
MP_REC_IMPL=tf32 MP_GEMM=tf32 timeout 300 python scripts/rtc_time.py > gpurun_out/rtc_time_tf32.log 2>&1; tail -8 gpurun_out/rtc_time_tf32.log
