# Full GPU visit: whole gpu suite, both bench arms, launch list and full-set captures of the three hot kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit $?"; cut -c1-200 gpurun_out/bench_cfg2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python scripts/prof_one.py --batch 256 --passes 2 --physics --tile 64 > gpurun_out/prof_list.log 2>&1; echo "ncu list exit $?"
# the launch list of the bench command itself (first 400 launches: warm-up + timed steps of the one-batch-at-a-time loop)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/prof_list_bench.log 2>&1; echo "ncu bench list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_tc -c 1 -o gpurun_out/prof_rec_tc_b256 python scripts/prof_one.py --batch 256 --passes 1 --tile 64 > gpurun_out/prof_a.log 2>&1; echo "ncu rec_tc exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -o gpurun_out/prof_gemm_tc python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_b.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_kernel -c 1 -o gpurun_out/prof_rec_b1 python scripts/prof_one.py --batch 1 --passes 1 > gpurun_out/prof_c.log 2>&1; echo "ncu rec b1 exit $?"
# (the FFMA cluster recurrence at B = 256, MP_REC_IMPL=ffma, is no default path any more: its capture stays in profiles/ from the earlier visits)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bias_act -s 1 -c 1 -o gpurun_out/prof_gemm_ffma2 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_f.log 2>&1; echo "ncu gemm ffma2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_h64_rows -c 1 -o gpurun_out/prof_rec_h64_rows python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_g.log 2>&1; echo "ncu h64 rows exit $?"
timeout 300 python scripts/time_gemm.py > gpurun_out/time_gemm.log 2>&1; cat gpurun_out/time_gemm.log
timeout 300 python scripts/time_eval.py > gpurun_out/time_eval.log 2>&1; cat gpurun_out/time_eval.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:physics_optimize -s 2 -c 1 -o gpurun_out/prof_k8_physics python scripts/time_physics.py --iters 1 > gpurun_out/prof_e.log 2>&1; echo "ncu k8 exit $?"
timeout 300 python scripts/time_physics.py > gpurun_out/time_physics.log 2>&1; tail -1 gpurun_out/time_physics.log
