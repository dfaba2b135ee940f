# Last visit of the round: the gpu suite on the final tree + the cfg4-shaped evaluate_pose timing.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 200 python scripts/time_evaluate.py > gpurun_out/time_evaluate.log 2>&1; echo "evaluate exit $?"; cat gpurun_out/time_evaluate.log
