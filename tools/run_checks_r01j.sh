# round-1 session-3 GPU checks (7): state-generation routes (auto by wave size, wide instantiation A/B)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_global_gpu.py tests/test_ocean_gpu.py tests/test_config1_gpu.py -x -q > gpurun_out/pytest_j1.log 2>&1; echo "pytest training rc=$?"; tail -4 gpurun_out/pytest_j1.log
for g in 3 6; do SML_TRAIN_SG_GROUP=$g timeout 300 python tools/bench_train_full.py --global-series --regions 384 > gpurun_out/train_sg$g.log 2> gpurun_out/train_sg.err; echo "group=$g rc=$? $(grep -o '"wall_s": [0-9.]*\|"stategen_s": [0-9.]*\|"gram_s": [0-9.]*' gpurun_out/train_sg$g.log | tr '\n' ' ')"; done
timeout 200 python tools/bench_train.py --regions 16 --cols 2000 --no-cublas > gpurun_out/bt_j16.log 2>&1; echo "wave 16: $(grep -o '"stategen_ms": [0-9.]*\|"gram_tflops_useful": [0-9.]*' gpurun_out/bt_j16.log | tr '\n' ' ')"
