# round-1 session-3 GPU checks (5): LU after the panel revert + ipiv staging; Gram pipeline depth A/B
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_train_gpu.py -x -q -k "mldivide or solver_paths" > gpurun_out/pytest_h1.log 2>&1; echo "pytest lu rc=$?"; tail -3 gpurun_out/pytest_h1.log
T="python tools/bench_train.py --regions 2 --cols 400 --solve --no-cublas"
SML_SOLVER=lu SML_LU_TIMING=1 timeout 300 $T > gpurun_out/bt_lu_h.log 2> gpurun_out/bt_lu_h.err; echo "bt lu rc=$?"; grep "sml lu" gpurun_out/bt_lu_h.err | head -2
SML_SOLVER=lu timeout 300 $T > gpurun_out/bt_lu_h2.log 2>> gpurun_out/bt_lu_h.err; echo "bt lu (no timing syncs) rc=$?"; grep -o '"solve_ms_per_region": [0-9.]*' gpurun_out/bt_lu_h2.log
G="python tools/bench_train.py --regions 16 --cols 2000 --no-cublas"
for st in 4 6 4 6; do SML_SYRK_STAGES=$st timeout 300 $G > gpurun_out/bt_st$st.log 2> gpurun_out/bt_st.err; echo "stages=$st rc=$? $(grep -o '"gram_tflops_useful": [0-9.]*\|"gram_ms": [0-9.]*' gpurun_out/bt_st$st.log | tr '\n' ' ')"; done
