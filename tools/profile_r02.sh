#!/bin/bash
# ncu evidence of round 2 (one B200; run under gpurun).  Every ncu command is preceded by the same
# command line without ncu.  Per-kernel event timing / ncu need plain launches: RTRG_NO_GRAPH=1.
set -x
mkdir -p gpurun_out
export RTRG_NO_GRAPH=1
BI="python tools/bench_integrals.py 64 1"
$BI > gpurun_out/plain_bi64.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bilinear -s 2 -c 1 -o gpurun_out/prof_bil_r02 $BI > gpurun_out/ncu_bi64.log 2>&1
B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-modes"
$B1 > gpurun_out/plain_b1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02.csv $B1 > gpurun_out/ncu_b1.log 2>&1
$B1 > gpurun_out/plain_b1b.log 2>&1 && \
ncu --set full --clock-control none -k regex:"k_attempt_local|k_stash|k_output|k_accept|k_growth_ode|k_assemble|k_extrap" -s 30 -c 14 -o gpurun_out/prof_hbm_r02 $B1 > gpurun_out/ncu_b1b.log 2>&1
B2="python bench.py --steps 1 --warmup 1 --mode full --cosmologies 256 --no-cpu-baseline --no-e2e --no-modes"
$B2 > gpurun_out/plain_b2.log 2>&1 && \
ncu --set full --clock-control none -k regex:"k_rhs|k_combine|k_final|k_pz" -s 40 -c 8 -o gpurun_out/prof_full_r02 $B2 > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv
