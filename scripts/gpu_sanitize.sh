# compute-sanitizer over small runs of every kernel family (memcheck + racecheck on the shared-memory heavy ones)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/time_physics.py --batch 3 --frames 12 --iters 1 > gpurun_out/san_k8_mem.log 2>&1; echo "k8 memcheck exit $?"; tail -3 gpurun_out/san_k8_mem.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/time_physics.py --batch 3 --frames 12 --iters 1 > gpurun_out/san_k8_race.log 2>&1; echo "k8 racecheck exit $?"; tail -3 gpurun_out/san_k8_race.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_one.py --batch 40 --frames 6 --passes 2 --physics > gpurun_out/san_net_mem.log 2>&1; echo "net memcheck exit $?"; tail -3 gpurun_out/san_net_mem.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_one.py --batch 1 --frames 9 --passes 2 > gpurun_out/san_b1_mem.log 2>&1; echo "b1 memcheck exit $?"; tail -3 gpurun_out/san_b1_mem.log
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/time_physics.py --batch 3 --frames 12 --iters 1 > gpurun_out/san_k8_race.log 2>&1; echo "k8 racecheck exit $?"; tail -2 gpurun_out/san_k8_race.log
MP_REC_IMPL=ffma timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_one.py --batch 40 --frames 6 --passes 2 > gpurun_out/san_net_ffma_mem.log 2>&1; echo "net ffma memcheck exit $?"; grep "^========= [A-Z]" gpurun_out/san_net_ffma_mem.log | sort | uniq -c | sort -rn | head -5
timeout 300 python -m pytest tests/test_gpu_physics.py -x -q 2>&1 | tail -2
