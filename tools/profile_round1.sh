set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train"
T="python tools/bench_train.py --regions 8 --cols 400 --solve --no-cublas"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $B > gpurun_out/ncu_b1.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_step|k_pack_grids|k_build_inputs|k_readout_finish" -s 8 -c 8 -o gpurun_out/prof_step_r01b -f $B > gpurun_out/ncu_b2.log 2>&1
echo rc=$?
$T > gpurun_out/plain_t.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_train_r01.csv $T > gpurun_out/ncu_t1.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_syrk_dmma" -c 1 -o gpurun_out/prof_syrk_r01b -f $T > gpurun_out/ncu_t2.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_chol_gemm" -s 58 -c 3 -o gpurun_out/prof_chol_r01 -f $T > gpurun_out/ncu_t3.log 2>&1
echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_chol_diag" -s 20 -c 1 -o gpurun_out/prof_choldiag_r01 -f $T > gpurun_out/ncu_t4.log 2>&1
echo rc=$?
tail -3 gpurun_out/ncu_b2.log gpurun_out/ncu_t3.log
ls -la gpurun_out/*.ncu-rep
