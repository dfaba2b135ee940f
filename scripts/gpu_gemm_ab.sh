# A/B visit for the FFMA GEMM tiles: gpu suite, isolated GEMM timings and the bench line under each switch.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python scripts/time_gemm.py > gpurun_out/gemm_new.log 2>&1; cat gpurun_out/gemm_new.log
MP_GEMM_FFMA2=0 timeout 200 python scripts/time_gemm.py > gpurun_out/gemm_noffma2.log 2>&1; cat gpurun_out/gemm_noffma2.log
MP_GEMM_WIDE=1 MP_GEMM_FFMA2=0 timeout 200 python scripts/time_gemm.py > gpurun_out/gemm_old.log 2>&1; cat gpurun_out/gemm_old.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_gemm_new.json 2> gpurun_out/bench_gemm_new.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench_gemm_new.json
MP_GEMM_FFMA2=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_gemm_noffma2.json 2>/dev/null; cut -c1-330 gpurun_out/bench_gemm_noffma2.json
MP_GEMM_WIDE=1 MP_GEMM_FFMA2=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_gemm_old.json 2>/dev/null; cut -c1-330 gpurun_out/bench_gemm_old.json
