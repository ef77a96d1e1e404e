#!/bin/bash
# A/B of the k_bilinear tuning variants (RTRG_BIL_VARIANT, kernels_integrals.cu: kBilVariants) on one B200:
#   tools/ab_bilinear.sh [B] [reps]  ->  gpurun_out/ab_bilinear.log
B=${1:-1024}; R=${2:-2}
mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6 7 8 9; do
  echo "== variant $v" >> gpurun_out/ab_bilinear.log
  RTRG_BIL_VARIANT=$v python tools/bench_integrals.py $B $R 2>&1 | grep -E "every product|output: P_T|z1l" >> gpurun_out/ab_bilinear.log
done
cat gpurun_out/ab_bilinear.log
