set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants and tc" > gpurun_out/pytest_rtc.log 2>&1; echo "rtc pytest exit $?"; tail -25 gpurun_out/pytest_rtc.log | cut -c1-300
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or cfg4" > gpurun_out/pytest_rtc2.log 2>&1; echo "rtc2 pytest exit $?"; tail -15 gpurun_out/pytest_rtc2.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rtc.json 2> gpurun_out/bench_rtc.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_rtc.json; tail -3 gpurun_out/bench_rtc.err
