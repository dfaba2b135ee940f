# projection GEMM with 8 epilogue warps + early accumulator hand-back, tensor-core linear1: parity suites, GEMM variants table, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_flat.py -q -x > gpurun_out/pytest_m.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/pytest_m.log | cut -c1-300 | tail -12
timeout 120 python scripts/gemm_il_ab.py > gpurun_out/gemm_variants.log 2>&1; grep -v FFMA gpurun_out/gemm_variants.log | cut -c1-120
timeout 100 python scripts/gemm_dbg.py 2>&1 | grep "gemm dbg" | awk "NR%2==0" > gpurun_out/gemm_dbg.log; cut -c1-260 gpurun_out/gemm_dbg.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.5 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench exit $?"; tail -3 gpurun_out/bench_m.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_m.json'))
print(d['value'], d['ms_per_step'], 'pinned', d['pinned_path']['value'], d['pinned_path']['ms_per_step'], 'e2e', d['e2e']['value'])
print('  ', {k:(round(v['ms_per_step'],3)) for k,v in d['kernels'].items()})
PY
