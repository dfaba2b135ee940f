set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300; grep "gemm M=" gpurun_out/pytest_gpu.log
timeout 600 python scripts/diag_precision.py 3000 > gpurun_out/diag_precision.log 2>&1; cat gpurun_out/diag_precision.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json | cut -c1-300; tail -5 gpurun_out/bench.err
