# A/B visit for the H = 64 row-per-thread recurrence + the evaluator kernels: gpu suite, bench under each switch.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h64_rows.json 2> gpurun_out/bench_h64_rows.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench_h64_rows.json; tail -3 gpurun_out/bench_h64_rows.err
MP_REC_H64=cluster timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h64_cluster.json 2>/dev/null; cut -c1-330 gpurun_out/bench_h64_cluster.json
