set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/time_physics.py --batch 3 --frames 12 --iters 1 > gpurun_out/san_k8_race.log 2>&1; echo "k8 racecheck exit $?"; tail -2 gpurun_out/san_k8_race.log
MP_REC_IMPL=ffma timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/prof_one.py --batch 40 --frames 6 --passes 2 > gpurun_out/san_net_ffma_mem.log 2>&1; echo "net ffma memcheck exit $?"; grep "^========= [A-Z]" gpurun_out/san_net_ffma_mem.log | sort | uniq -c | sort -rn | head -5
timeout 300 python -m pytest tests/test_gpu_physics.py -x -q 2>&1 | tail -2
