# compute-sanitizer over the kernels added after gpu_sanitize.sh: the 128-sequence recurrence (lstm_rec_f16w.cu, TMA layer output) and
# the CTA-pair projection GEMM (gemm_f16.cu, cta_group::2), each through a pipeline slot with the wide tile (batch 128 x 4 frames:
# M = 512 rows would stay below the tensor-core GEMM's threshold, MP_GEMM=tc lifts it).
set -x
mkdir -p gpurun_out
for tool in racecheck initcheck synccheck memcheck; do
  MP_GEMM=tc MP_GEMM_PAIR=1 timeout 600 compute-sanitizer --tool $tool --print-limit 2000 python scripts/prof_one.py --batch 128 --frames 4 --passes 1 --tile 128 > gpurun_out/san_wide_pair_$tool.log 2>&1; echo "wide + pair $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|^done" gpurun_out/san_wide_pair_$tool.log | sort | uniq -c | sort -rn | head -6
  grep -o "Device Frame: void mp::<unnamed>::[a-z_0-9]*" gpurun_out/san_wide_pair_$tool.log | sort | uniq -c
done
