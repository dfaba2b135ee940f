# A/B of two builds of the library in one visit: default vs mobileposer_b200/lib/libmp_expf.so (swapped in place)
set -x
mkdir -p gpurun_out
L=mobileposer_b200/lib
for round in 1 2; do
  bash scripts/gpu_rtc_ts.sh 2>&1 | grep "lstm_rec_tc_h256\|s=6"
  cp $L/libmobileposer_b200.so $L/tmp.so; cp $L/libmp_expf.so $L/libmobileposer_b200.so
  echo "--- variant B (expf)"; bash scripts/gpu_rtc_ts.sh 2>&1 | grep "lstm_rec_tc_h256\|s=6"
  cp $L/tmp.so $L/libmobileposer_b200.so
  echo "--- variant A"
done
