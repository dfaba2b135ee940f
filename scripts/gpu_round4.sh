set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json | cut -c1-400; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python scripts/prof_one.py --batch 256 --passes 2 > gpurun_out/prof_list.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec -c 1 -o gpurun_out/prof_rec_b256 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_b256.log 2>&1; echo "ncu full b256 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -o gpurun_out/prof_gemm_tc python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_gemm.log 2>&1; echo "ncu full gemm exit $?"
