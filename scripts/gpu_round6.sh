set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json | cut -c1-200; tail -5 gpurun_out/bench.err
MP_REC_CLUSTER=16 timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg2_c16.json 2> gpurun_out/bench_cfg2_c16.err; echo "c16 exit $?"; cat gpurun_out/bench_cfg2_c16.json | cut -c1-300; tail -3 gpurun_out/bench_cfg2_c16.err
timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit $?"; cat gpurun_out/bench_cfg2.json | cut -c1-300
