set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
MP_REC_IMPL=simple timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_simple.log 2>&1; echo "smoke simple exit $?"; tail -5 gpurun_out/smoke_simple.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gpu.log
