# Final visit of the round: whole gpu suite, smoke, both bench arms, cfg2, launch lists and full-set captures of the two kernels that
# changed last (the wide recurrence with TMA output / TMA gate pre-activations, the CTA-pair projection at K = 512).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -6
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches.csv python scripts/prof_one.py --batch 256 --passes 2 --physics --tile 128 > gpurun_out/prof_list.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg4 --min-seconds 0 > gpurun_out/prof_list_bench.log 2>&1; echo "ncu bench list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_f16w -c 1 -o gpurun_out/prof_rec_f16w_b256 python scripts/prof_one.py --batch 256 --passes 1 --tile 128 > gpurun_out/prof_a.log 2>&1; echo "ncu rec_f16w exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 2 -c 1 -o gpurun_out/prof_gemm_f16_k512 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_b.log 2>&1; echo "ncu gemm exit $?"
timeout 200 python scripts/rec_wide_ab.py > gpurun_out/rec_wide_ab.log 2>&1; tail -2 gpurun_out/rec_wide_ab.log
timeout 100 python scripts/recw_ts.py 2>&1 | grep "recw ts" | head -5 > gpurun_out/recw_ts.log
