# round-1 session-3 GPU checks (6): in-kernel time loop of the training state generation (k_train_stategen), warp-parallel interchanges
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_global_gpu.py tests/test_ocean_gpu.py tests/test_config1_gpu.py -x -q > gpurun_out/pytest_i1.log 2>&1; echo "pytest training rc=$?"; tail -4 gpurun_out/pytest_i1.log
timeout 300 python tools/bench_train_full.py --global-series > gpurun_out/train_full_i_on.log 2> gpurun_out/train_full_i.err; echo "train full overlap rc=$?"; cat gpurun_out/train_full_i_on.log
timeout 300 python tools/bench_train_full.py --global-series --no-overlap > gpurun_out/train_full_i_off.log 2>> gpurun_out/train_full_i.err; echo "train full serial rc=$?"; cat gpurun_out/train_full_i_off.log
T="python tools/bench_train.py --regions 2 --cols 400 --solve --no-cublas"
SML_SOLVER=lu SML_LU_TIMING=1 timeout 300 $T > gpurun_out/bt_lu_i.log 2> gpurun_out/bt_lu_i.err; echo "bt lu rc=$?"; grep "sml lu" gpurun_out/bt_lu_i.err | head -2
SML_SOLVER=lu timeout 300 $T > gpurun_out/bt_lu_i2.log 2>> gpurun_out/bt_lu_i.err; echo "bt lu (no timing syncs) rc=$?"; grep -o '"solve_ms_per_region": [0-9.]*' gpurun_out/bt_lu_i2.log
