# round-1 session-3 GPU checks (3): DMMA trailing update of the LU, refreshed config-5 sweep, ncu capture of k_update_sx
mkdir -p gpurun_out
SML_LU_TIMING=1 timeout 400 python -m pytest tests/test_train_gpu.py tests/test_config1_gpu.py -x -q > gpurun_out/pytest_f1.log 2>&1; echo "pytest train+config1 rc=$?"; tail -4 gpurun_out/pytest_f1.log
SML_SOLVER=lu SML_LU_TIMING=1 timeout 300 python tools/bench_train.py --regions 4 --cols 400 --solve --no-cublas > gpurun_out/bt_lu_f.log 2> gpurun_out/bt_lu_f.err; echo "bt lu rc=$?"; cat gpurun_out/bt_lu_f.log; grep "sml lu" gpurun_out/bt_lu_f.err | head -4
SML_SOLVER=lu timeout 300 python tools/bench_train.py --regions 4 --cols 400 --solve --no-cublas > gpurun_out/bt_lu_f2.log 2>> gpurun_out/bt_lu_f.err; echo "bt lu (no timing syncs) rc=$?"; cat gpurun_out/bt_lu_f2.log
S="python tools/sweep.py --m 6000 --deg 6 --steps 3"
timeout 200 $S > gpurun_out/plain_s_f.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_update_sx" -s 3 -c 1 -o gpurun_out/prof_update_sx_r01d -f $S > gpurun_out/ncu_s_f.log 2>&1
echo "ncu update_sx rc=$?"
timeout 900 python tools/sweep.py --md gpurun_out/sweep_r01d.md > gpurun_out/sweep_r01d.jsonl 2> gpurun_out/sweep_f.err; echo "sweep rc=$?"; cat gpurun_out/sweep_r01d.md
