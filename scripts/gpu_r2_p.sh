# wide (128-sequence) recurrence in the pipelined step: bench A/B via MP_REC_WIDE
set -x
mkdir -p gpurun_out
for v in 1 0; do
  MP_REC_WIDE=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1.5 > gpurun_out/bench_p_$v.json 2> gpurun_out/bench_p_$v.err; echo "bench exit $?"; tail -2 gpurun_out/bench_p_$v.err
done
python - <<'PY'
import json
for n in ('1','0'):
    d=json.load(open(f'gpurun_out/bench_p_{n}.json'))
    print('wide' if n=='1' else 'n64 ', d['value'], d['ms_per_step'], 'pinned', d['pinned_path']['value'], d['pinned_path']['ms_per_step'], 'e2e', d['e2e']['value'], 'one at a time', d['one_batch_at_a_time']['ms_per_step'])
    print('  ', {k:(round(v['ms_per_step'],3)) for k,v in d['kernels'].items()})
PY
