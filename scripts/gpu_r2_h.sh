set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR|narrow" gpurun_out/pytest_gpu.log | cut -c1-200 | tail -14
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 200 python scripts/rtc_time.py 2>&1 | grep -E "lstm_rec|gemm|split|linear"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_h.json; tail -5 gpurun_out/bench_h.err
MP_LINEAR2_FFMA=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1 > gpurun_out/bench_h_l2ffma.json 2> gpurun_out/bench_h_l2ffma.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_h_l2ffma.json
