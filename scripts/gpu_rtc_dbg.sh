set -x
mkdir -p gpurun_out
for nb in 64 52 32; do
MP_REC_NB=$nb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nb$nb.json 2> gpurun_out/bench_nb$nb.err; echo "nb=$nb exit $?"; python -c "
import json;b=json.load(open('gpurun_out/bench_nb$nb.json'));print('NB=$nb value',round(b['value']),'ms',round(b['ms_per_step'],2),{k:round(v['ms_per_step'],2) for k,v in b['kernels'].items()})"
done
