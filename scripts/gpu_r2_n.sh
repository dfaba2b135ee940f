# training loop: GPU tests, step timing; pair-kernel recheck after the relaxed arrives
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -rP > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; grep -E "^\[train\]|passed|failed|^E  |Error" gpurun_out/pytest_train.log | cut -c1-330 | tail -20
timeout 600 python scripts/time_train.py > gpurun_out/time_train.log 2>&1; grep "train time" gpurun_out/time_train.log || tail -5 gpurun_out/time_train.log
timeout 120 python scripts/gemm_il_ab.py 2>&1 | grep "il=0" | cut -c1-100
