#!/bin/bash
# ncu evidence of round 1 (run on a B200 through gpurun; one ncu tool use per call).  Each profiled command first runs
# plain and must exit 0.  Outputs land in gpurun_out/; tools/ncu_extract.py turns the .ncu-rep files into profiles/*.csv.
set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train"
T="python tools/bench_train.py --regions 8 --cols 400 --solve --no-cublas"
S="python tools/sweep.py --m 6000 --deg 6 --steps 3"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv $B > gpurun_out/ncu_b1.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_step|k_pack_grids|k_build_inputs|k_readout_finish" -s 8 -c 8 -o gpurun_out/prof_step_r01c -f $B > gpurun_out/ncu_b2.log 2>&1
echo rc=$?
$S > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_update" -s 3 -c 1 -o gpurun_out/prof_update_r01c -f $S > gpurun_out/ncu_s1.log 2>&1
echo rc=$?
$T > gpurun_out/plain_t.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train_r01c.csv $T > gpurun_out/ncu_t1.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_syrk_dmma" -c 1 -o gpurun_out/prof_syrk_r01c -f $T > gpurun_out/ncu_t2.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_chol_gemm" -s 58 -c 3 -o gpurun_out/prof_chol_r01c -f $T > gpurun_out/ncu_t3.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_chol_diag" -s 20 -c 1 -o gpurun_out/prof_choldiag_r01c -f $T > gpurun_out/ncu_t4.log 2>&1
echo rc=$?
ls -la gpurun_out/*r01c*
