set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tensor_core" > gpurun_out/pytest_tc.log 2>&1; echo "tc pytest exit $?"; tail -15 gpurun_out/pytest_tc.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json | cut -c1-1500; tail -5 gpurun_out/bench.err
