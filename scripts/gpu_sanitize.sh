# compute-sanitizer: (1) the minimal DSMEM bulk-copy reproducer (three variants) under all four tools; (2) all four tools over the
# fp16-split tensor-core recurrence ALONE (whole net at batch 40: no FFMA cluster kernel in the process), the FFMA throughput cluster
# kernel alone, and the FFMA latency kernel (batch 1).
set -x
mkdir -p gpurun_out
cd scripts/sanitizer && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o dsmem_bulk_repro dsmem_bulk_repro.cu && cd ../..
./scripts/sanitizer/dsmem_bulk_repro
for tool in memcheck racecheck initcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool ./scripts/sanitizer/dsmem_bulk_repro > gpurun_out/san_repro_$tool.log 2>&1; echo "repro $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|dsmem bulk|not located" gpurun_out/san_repro_$tool.log | sort | uniq -c | sort -rn | head -8
done
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 3000 python scripts/prof_one.py --batch 40 --frames 5 --passes 1 > gpurun_out/san_net_f16_$tool.log 2>&1; echo "net (f16 recurrence) $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|^done" gpurun_out/san_net_f16_$tool.log | sort | uniq -c | sort -rn | head -6
  grep -o "Device Frame: void mp::<unnamed>::[a-z_0-9]*" gpurun_out/san_net_f16_$tool.log | sort | uniq -c
  MP_REC_IMPL=ffma timeout 600 compute-sanitizer --tool $tool --print-limit 3000 python scripts/prof_one.py --batch 40 --frames 5 --passes 1 > gpurun_out/san_net_ffma_$tool.log 2>&1; echo "net (FFMA cluster recurrence) $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|^done" gpurun_out/san_net_ffma_$tool.log | sort | uniq -c | sort -rn | head -6
  grep -o "Device Frame: void mp::<unnamed>::[a-z_0-9]*" gpurun_out/san_net_ffma_$tool.log | sort | uniq -c
  timeout 600 compute-sanitizer --tool $tool python scripts/prof_one.py --batch 1 --frames 6 --passes 1 > gpurun_out/san_rec_b1_$tool.log 2>&1; echo "rec b1 $tool exit $?"
  grep -E "^========= [A-Z]|ERROR SUMMARY|RACECHECK SUMMARY|^done" gpurun_out/san_rec_b1_$tool.log | sort | uniq -c | sort -rn | head -6
done
