set -x
mkdir -p gpurun_out
MP_RTC_DBG=0 timeout 60 python scripts/rtc_debug.py 4 3 > gpurun_out/dbg_a.log 2>&1; echo "T3 exit $?"; grep -E "max \|tc|CUDA error|Error" gpurun_out/dbg_a.log | head -2 | cut -c1-200
MP_RTC_DBG=0 timeout 60 python scripts/rtc_debug.py 40 50 > gpurun_out/dbg_b.log 2>&1; echo "B40 exit $?"; grep -E "max \|tc|CUDA error|Error" gpurun_out/dbg_b.log | head -2 | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/pytest_rtc.log 2>&1; echo "variants exit $?"; tail -6 gpurun_out/pytest_rtc.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or cfg4 or float64" > gpurun_out/pytest_rtc2.log 2>&1; echo "rtc2 pytest exit $?"; tail -8 gpurun_out/pytest_rtc2.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rtc.json 2> gpurun_out/bench_rtc.err; echo "bench exit $?"; cut -c1-260 gpurun_out/bench_rtc.json; tail -3 gpurun_out/bench_rtc.err
