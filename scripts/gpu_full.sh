# Full GPU visit: whole gpu suite, smoke, both bench arms, launch lists and full-set captures of the hot kernels
# (python scripts/summarize_profile.py r02 then refreshes profiles/).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -6
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --workload cfg2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit $?"; cut -c1-200 gpurun_out/bench_cfg2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches.csv python scripts/prof_one.py --batch 256 --passes 2 --physics --tile 128 > gpurun_out/prof_list.log 2>&1; echo "ncu list exit $?"
# the launch list of the bench command itself (first 400 launches: warm-up + timed steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg4 --min-seconds 0 > gpurun_out/prof_list_bench.log 2>&1; echo "ncu bench list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_f16w -c 1 -o gpurun_out/prof_rec_f16w_b256 python scripts/prof_one.py --batch 256 --passes 1 --tile 128 > gpurun_out/prof_a.log 2>&1; echo "ncu rec_f16w exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_f16_kernel -c 1 -o gpurun_out/prof_rec_f16_b256 python scripts/prof_one.py --batch 256 --passes 1 --tile 64 > gpurun_out/prof_a2.log 2>&1; echo "ncu rec_f16 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 1 -c 1 -o gpurun_out/prof_gemm_f16 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_b.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_kernel -s 2 -c 1 -o gpurun_out/prof_rec_b1 python scripts/prof_one.py --batch 1 --passes 1 > gpurun_out/prof_c.log 2>&1; echo "ncu rec b1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bias_act -s 1 -c 1 -o gpurun_out/prof_gemm_ffma2 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_f.log 2>&1; echo "ncu gemm ffma2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_h64_rows -c 1 -o gpurun_out/prof_rec_h64_rows python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_g.log 2>&1; echo "ncu h64 rows exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:physics_optimize -s 2 -c 1 -o gpurun_out/prof_k8_physics python scripts/time_physics.py --iters 1 > gpurun_out/prof_e.log 2>&1; echo "ncu k8 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 0 -c 1 -o gpurun_out/prof_gemm_f16_linear1 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_b2.log 2>&1; echo "ncu gemm linear1 exit $?"
timeout 300 python scripts/time_gemm16.py > gpurun_out/time_gemm16.log 2>&1; cat gpurun_out/time_gemm16.log | grep gemm
timeout 300 python scripts/time_train.py > gpurun_out/time_train.log 2>&1; grep "train time" gpurun_out/time_train.log
timeout 200 python scripts/rec_wide_ab.py > gpurun_out/rec_wide_ab.log 2>&1; tail -2 gpurun_out/rec_wide_ab.log
timeout 100 python scripts/recw_ts.py 2>&1 | grep "recw ts" | head -5 > gpurun_out/recw_ts.log
timeout 300 python scripts/time_physics.py > gpurun_out/time_physics.log 2>&1; tail -1 gpurun_out/time_physics.log
timeout 200 python scripts/rtc_time.py > gpurun_out/rtc_time.log 2>&1; cat gpurun_out/rtc_time.log | tail -6
MP_RTC_TS=1 timeout 100 python scripts/rtc_debug.py 256 40 f16 > gpurun_out/rtc_ts.log 2>&1; grep "rtc ts" gpurun_out/rtc_ts.log | sed -n 2,7p
