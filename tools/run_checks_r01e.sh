# round-1 session-3 GPU checks (2): TMA-staged update kernel (default), cooperative LU panel, rolling average
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_rolling_average.py tests/test_ocean_gpu.py -x -q > gpurun_out/pytest_e1.log 2>&1; echo "pytest engine rc=$?"; tail -3 gpurun_out/pytest_e1.log
SML_LU_TIMING=1 timeout 400 python -m pytest tests/test_train_gpu.py -x -q -k "mldivide or solver_paths" > gpurun_out/pytest_e2.log 2>&1; echo "pytest lu rc=$?"; tail -5 gpurun_out/pytest_e2.log
SML_SOLVER=lu SML_LU_TIMING=1 timeout 300 python tools/bench_train.py --regions 4 --cols 400 --solve --no-cublas > gpurun_out/bt_lu_e.log 2> gpurun_out/bt_lu_e.err; echo "bt lu rc=$?"; cat gpurun_out/bt_lu_e.log; grep "sml lu" gpurun_out/bt_lu_e.err | head -4
timeout 200 python tools/ab_update.py > gpurun_out/ab_update_e6.jsonl 2> gpurun_out/ab_update.err; echo "ab rc=$?"; cut -c1-175 gpurun_out/ab_update_e6.jsonl
timeout 200 python tools/ab_update.py --deg 24 > gpurun_out/ab_update_e24.jsonl 2>> gpurun_out/ab_update.err; echo "ab24 rc=$?"; cut -c1-175 gpurun_out/ab_update_e24.jsonl
timeout 200 python tools/ab_update.py --regions 144 > gpurun_out/ab_update_e144.jsonl 2>> gpurun_out/ab_update.err; echo "ab144 rc=$?"; cut -c1-175 gpurun_out/ab_update_e144.jsonl
