# compute-sanitizer: (1) the minimal DSMEM bulk-copy reproducer under all four tools; (2) racecheck / initcheck / synccheck of both
# cluster recurrences (fp16-split tensor-core kernel, FFMA latency kernel) on tiny runs; (3) memcheck of the same for the record.
set -x
mkdir -p gpurun_out
cd scripts/sanitizer && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o dsmem_bulk_repro dsmem_bulk_repro.cu && cd ../..
./scripts/sanitizer/dsmem_bulk_repro
for tool in memcheck racecheck initcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool ./scripts/sanitizer/dsmem_bulk_repro > gpurun_out/san_repro_$tool.log 2>&1; echo "repro $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|dsmem bulk" gpurun_out/san_repro_$tool.log | sort | uniq -c | sort -rn | head -6
done
for tool in racecheck initcheck synccheck memcheck; do
  MP_REC_IMPL=f16 timeout 600 compute-sanitizer --tool $tool python scripts/rtc_debug.py 20 5 f16 > gpurun_out/san_rec_f16_$tool.log 2>&1; echo "rec f16 $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|max \|tc" gpurun_out/san_rec_f16_$tool.log | sort | uniq -c | sort -rn | head -6
  timeout 600 compute-sanitizer --tool $tool python scripts/prof_one.py --batch 1 --frames 6 --passes 1 > gpurun_out/san_rec_b1_$tool.log 2>&1; echo "rec b1 $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san_rec_b1_$tool.log | sort | uniq -c | sort -rn | head -6
done
