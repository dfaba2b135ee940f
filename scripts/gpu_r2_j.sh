set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-200 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 --min-seconds 1 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench exit $?"; tail -3 gpurun_out/bench_j.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_j.json').read())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e compact', round(d['e2e_compact']['value']), '| pinned', round(d['pinned_path']['value']), round(d['pinned_path']['e2e']['value']), round(d['pinned_path']['e2e_compact']['value']))"
# latency kernel: source-level stall profile at batch 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_kernel -s 2 -c 1 -o gpurun_out/prof_rec_b1 python scripts/prof_one.py --batch 1 --passes 1 > gpurun_out/prof_c.log 2>&1; echo "ncu rec b1 exit $?"
