# linear1 on the fp16-split tensor-core kernel: parity suites, then A/B bench against the FFMA linear1
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_flat.py -q -x > gpurun_out/pytest_l.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/pytest_l.log | cut -c1-300 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.5 > gpurun_out/bench_l1tc.json 2> gpurun_out/bench_l1tc.err; echo "bench exit $?"; cut -c1-260 gpurun_out/bench_l1tc.json; tail -3 gpurun_out/bench_l1tc.err
MP_LINEAR1_FFMA=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.5 > gpurun_out/bench_l1ffma.json 2> gpurun_out/bench_l1ffma.err; echo "bench exit $?"; cut -c1-260 gpurun_out/bench_l1ffma.json
python - <<'PY'
import json
for n in ('l1tc','l1ffma'):
    d=json.load(open(f'gpurun_out/bench_{n}.json'))
    print(n, d['value'], d['ms_per_step'], 'pinned', d['pinned_path']['value'], d['pinned_path']['ms_per_step'])
    print('  ', {k:(round(v['ms_per_step'],3)) for k,v in d['kernels'].items()})
PY
