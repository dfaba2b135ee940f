set -x
mkdir -p gpurun_out
for d in 4 8 10; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.0 --depth $d > gpurun_out/bench_d$d.json 2> gpurun_out/bench_d$d.err; echo "bench exit $?"
done
python - <<'PY'
import json
for n in ('4','8','10'):
    d=json.load(open(f'gpurun_out/bench_d{n}.json'))
    print('depth', n, d['value'], d['ms_per_step'], 'pinned', d['pinned_path']['ms_per_step'], 'e2e', d['e2e']['value'])
PY
