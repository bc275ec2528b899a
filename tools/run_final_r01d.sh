# round-1 session-3 final GPU pass: full GPU test suite, default bench, whole-model training with the overlapped schedule,
# ncu launch lists (bench step, LU route), sanitizer passes over the kernels added in this session
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01d.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/pytest_gpu_r01d.log
timeout 600 python bench.py > gpurun_out/bench_r01d.log 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_r01d.log
timeout 300 python tools/bench_train_full.py --global-series > gpurun_out/train_full_r01d.log 2> gpurun_out/train_full_r01d.err; echo "train full rc=$?"; cat gpurun_out/train_full_r01d.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv $B > gpurun_out/ncu_b_d.log 2>&1; echo "ncu bench rc=$?"
T="python tools/bench_train.py --regions 2 --cols 400 --solve --no-cublas"
SML_SOLVER=lu timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_lu_r01d.csv $T > gpurun_out/ncu_lu_d.log 2>&1; echo "ncu lu rc=$?"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_rolling_average.py tests/test_train_gpu.py tests/test_engine_gpu.py -m gpu -x -q -k "rolling or mldivide or overlapped or synchronize" > gpurun_out/memcheck_r01d.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_r01d.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_train_gpu.py tests/test_engine_gpu.py -m gpu -x -q -k "mldivide or synchronize" > gpurun_out/racecheck_r01d.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_r01d.log
