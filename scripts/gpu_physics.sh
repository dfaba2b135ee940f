set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_physics.py -x -q -s > gpurun_out/pytest_physics.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_physics.log | cut -c1-300
timeout 300 python scripts/time_physics.py > gpurun_out/time_physics.log 2>&1; cat gpurun_out/time_physics.log | tail -3
timeout 300 python scripts/time_physics.py --batch 1 >> gpurun_out/time_physics.log 2>&1; tail -1 gpurun_out/time_physics.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:physics_optimize -s 2 -c 1 -o gpurun_out/prof_physics python scripts/time_physics.py --iters 1 > gpurun_out/prof_p.log 2>&1; echo "ncu exit $?"
