# CTA-pair projection kernel as the default for K >= 256: parity suites, bench A/B against MP_GEMM_PAIR=0
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_flat.py -q -x > gpurun_out/pytest_o.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/pytest_o.log | cut -c1-300 | tail -12
for v in pair single; do
  if [ $v = single ]; then export MP_GEMM_PAIR=0; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.5 > gpurun_out/bench_o_$v.json 2> gpurun_out/bench_o_$v.err; echo "bench exit $?"; tail -2 gpurun_out/bench_o_$v.err
done
python - <<'PY'
import json
for n in ('pair','single'):
    d=json.load(open(f'gpurun_out/bench_o_{n}.json'))
    print(n, d['value'], d['ms_per_step'], 'pinned', d['pinned_path']['value'], d['pinned_path']['ms_per_step'], 'e2e', d['e2e']['value'])
    print('  ', {k:(round(v['ms_per_step'],3)) for k,v in d['kernels'].items()})
PY
