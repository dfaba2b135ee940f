set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "wide or pipelined_slots or tile_policy" > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/pytest_q.log | cut -c1-300 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.5 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench exit $?"; tail -2 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print(d['value'], d['ms_per_step'], 'pinned', d['pinned_path']['value'], d['pinned_path']['ms_per_step'], 'e2e', d['e2e']['value'], 'one at a time', d['one_batch_at_a_time']['ms_per_step'], d['e2e']['one_batch_at_a_time'])
print('  ', {k:(round(v['ms_per_step'],3)) for k,v in d['kernels'].items()})
print(d['roofline'])
PY
