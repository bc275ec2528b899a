# round-1 session-3 GPU checks: overlapped training schedule, shared-memory update kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_train_gpu.py tests/test_train_global_gpu.py tests/test_ocean_gpu.py -x -q > gpurun_out/pytest_d1.log 2>&1; echo "pytest train rc=$?"; tail -3 gpurun_out/pytest_d1.log
SML_UPDATE_KERNEL=sx timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_ocean_gpu.py -x -q > gpurun_out/pytest_d2.log 2>&1; echo "pytest sx rc=$?"; tail -3 gpurun_out/pytest_d2.log
timeout 200 python tools/ab_update.py > gpurun_out/ab_update_d6.jsonl 2> gpurun_out/ab_update.err; echo "ab rc=$?"; cut -c1-200 gpurun_out/ab_update_d6.jsonl
timeout 200 python tools/ab_update.py --deg 24 > gpurun_out/ab_update_d24.jsonl 2>> gpurun_out/ab_update.err; echo "ab24 rc=$?"; cut -c1-200 gpurun_out/ab_update_d24.jsonl
timeout 200 python tools/ab_update.py --regions 144 > gpurun_out/ab_update_r144.jsonl 2>> gpurun_out/ab_update.err; echo "ab144 rc=$?"; cut -c1-200 gpurun_out/ab_update_r144.jsonl
timeout 300 python tools/bench_train_full.py --regions 192 --wave 96 --global-series > gpurun_out/train_ovl_on.log 2> gpurun_out/train_ovl.err; echo "ovl on rc=$?"; cat gpurun_out/train_ovl_on.log
timeout 300 python tools/bench_train_full.py --regions 192 --wave 96 --global-series --no-overlap > gpurun_out/train_ovl_off.log 2>> gpurun_out/train_ovl.err; echo "ovl off rc=$?"; cat gpurun_out/train_ovl_off.log
