set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python scripts/diag_precision.py 3000 > gpurun_out/diag_precision.log 2>&1; cat gpurun_out/diag_precision.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec -c 2 -o gpurun_out/prof_rec_b256 python scripts/prof_one.py --batch 256 --passes 1 > gpurun_out/prof_b256.log 2>&1; echo "ncu full b256 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec -c 2 -o gpurun_out/prof_rec_b1 python scripts/prof_one.py --batch 1 --passes 1 > gpurun_out/prof_b1.log 2>&1; echo "ncu full b1 exit $?"
