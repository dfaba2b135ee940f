# compute-sanitizer over small runs of every kernel family (memcheck + racecheck on the shared-memory heavy ones)
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/time_physics.py --batch 3 --frames 12 --iters 1 > gpurun_out/san_k8_mem.log 2>&1; echo "k8 memcheck exit $?"; tail -3 gpurun_out/san_k8_mem.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/time_physics.py --batch 3 --frames 12 --iters 1 > gpurun_out/san_k8_race.log 2>&1; echo "k8 racecheck exit $?"; tail -3 gpurun_out/san_k8_race.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_one.py --batch 40 --frames 6 --passes 2 --physics > gpurun_out/san_net_mem.log 2>&1; echo "net memcheck exit $?"; tail -3 gpurun_out/san_net_mem.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_one.py --batch 1 --frames 9 --passes 2 > gpurun_out/san_b1_mem.log 2>&1; echo "b1 memcheck exit $?"; tail -3 gpurun_out/san_b1_mem.log
