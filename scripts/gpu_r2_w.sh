set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_evaluate.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "cfg4 exit $?"; tail -2 gpurun_out/bench_cfg4_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cfg4_n1.json'))
print('cfg4 n1', d['value'], d['ms_per_step'], d.get('table_sha16_rounded_1e-2'), d.get('all_gather_shape'))
PY
