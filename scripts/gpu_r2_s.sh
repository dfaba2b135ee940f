set -x
mkdir -p gpurun_out
for w in 8 12 2; do
  MP_K8_WARPS=$w timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.0 > gpurun_out/bench_k8w_$w.json 2> gpurun_out/bench_k8w_$w.err; echo "bench exit $?"
done
python - <<'PY'
import json
for n in ('8','12','2'):
    d=json.load(open(f'gpurun_out/bench_k8w_{n}.json'))
    print('K8 warps', n, d['value'], d['ms_per_step'], 'pinned', d['pinned_path']['ms_per_step'], 'k8', d['kernels']['k8_physics']['ms_per_step'])
PY
