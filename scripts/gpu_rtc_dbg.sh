set -x
mkdir -p gpurun_out
MP_RTC_TS=1 MP_REC_IMPL=tc timeout 120 python scripts/rtc_debug.py 256 20 > gpurun_out/ts.log 2>&1; echo "ts exit $?"; grep -E "rtc ts|max" gpurun_out/ts.log | sed -n '1,4p;$p' | cut -c1-250
MP_REC_IMPL=tc timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or ragged or cfg3 or cfg4 or float64 or batch_equals" > gpurun_out/pytest_rtc.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/pytest_rtc.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rtc.json 2> gpurun_out/bench_rtc.err; echo "bench exit $?"; python -c "
import json;b=json.load(open('gpurun_out/bench_rtc.json'));print('value',round(b['value']),'ms',round(b['ms_per_step'],2),'e2e',round(b['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in b['kernels'].items()})"; tail -3 gpurun_out/bench_rtc.err
