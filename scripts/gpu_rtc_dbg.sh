set -x
mkdir -p gpurun_out
for nb in 16 32 64; do
MP_RTC_TS=1 MP_REC_IMPL=tc MP_REC_NB=$nb timeout 120 python scripts/rtc_debug.py 256 12 > gpurun_out/ts_$nb.log 2>&1; echo "ts nb=$nb exit $?"; grep -E "rtc ts|max" gpurun_out/ts_$nb.log | sed -n '1p;3p;$p' | cut -c1-330
done
