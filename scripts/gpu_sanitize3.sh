set -x
mkdir -p gpurun_out
for tool in racecheck memcheck; do
  timeout 200 compute-sanitizer --tool $tool --print-limit 200 python scripts/san_gemm_pair.py > gpurun_out/san_gemm_pair_$tool.log 2>&1; echo "pair gemm $tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|pair gemm max" gpurun_out/san_gemm_pair_$tool.log
  grep "Race reported" gpurun_out/san_gemm_pair_$tool.log | grep -o "gemm_f16.cu:[0-9]*" | sort | uniq -c
done
MP_GEMM=ffma timeout 240 compute-sanitizer --tool initcheck --print-limit 50 python scripts/prof_one.py --batch 128 --frames 4 --passes 1 --tile 128 > gpurun_out/san_wide_ffma_initcheck.log 2>&1; echo "wide (FFMA gemm) initcheck exit $?"
grep -E "ERROR SUMMARY|^done" gpurun_out/san_wide_ffma_initcheck.log; grep -o "Device Frame: void mp::<unnamed>::[a-z_0-9]*" gpurun_out/san_wide_ffma_initcheck.log | sort | uniq -c
timeout 120 python scripts/gemm_il_ab.py 2>&1 | grep "il=0" | cut -c1-100
