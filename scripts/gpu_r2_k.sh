set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -rP > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; grep -E "^\[train\]|passed|failed|^E  |Error" gpurun_out/pytest_train.log | cut -c1-250 | tail -20
