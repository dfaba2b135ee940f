# round 2, visit C: A/B of the fp16 recurrence variants after the epilogue reorder (exchange before global memory traffic)
set -x
mkdir -p gpurun_out
for v in "" "MP_RF16_ACT=fast" "MP_RF16_EXCHANGE=direct" "MP_RF16_EXCHANGE=direct MP_RF16_ACT=fast" "MP_RF16_EXCHANGE=direct MP_RF16_ACT=fast MP_RF16_RAGGED=1"; do
  echo "== variant: $v"
  env $v timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | tail -1
  env $v timeout 100 python scripts/rtc_debug.py 70 24 f16 2>&1 | tail -1
  env $v timeout 200 python scripts/rtc_time.py 2>&1 | grep lstm_rec
  env $v MP_RTC_TS=1 timeout 100 python scripts/rtc_debug.py 256 40 f16 2>&1 | grep "rtc ts" | sed -n 2,4p
done
