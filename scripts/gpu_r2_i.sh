set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_physics.py tests/test_gpu_evaluate.py -q > gpurun_out/pytest_phys.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_phys.log | cut -c1-200
for w in 1 4 8 12; do
  MP_K8_WARPS=$w timeout 300 python scripts/time_physics.py 2>&1 | tail -1
  MP_K8_WARPS=$w timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1 > gpurun_out/bench_i_w$w.json 2> gpurun_out/bench_i_w$w.err; echo "bench w=$w exit $?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_i_w$w.json').read())
print('K8 warps/CTA $w: value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'one-at-a-time ms', round(d['one_batch_at_a_time']['ms_per_step'],3), '| pinned', round(d['pinned_path']['value']), round(d['pinned_path']['ms_per_step'],3), '| k8 ms', round(d['kernels']['k8_physics']['ms_per_step'],3))"
done
