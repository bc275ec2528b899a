# round-1 session-3 GPU checks (4): LU with the two-barrier panel, parallel interchanges and the DMMA trailing update
mkdir -p gpurun_out
SML_LU_TIMING=1 timeout 400 python -m pytest tests/test_train_gpu.py tests/test_config1_gpu.py -x -q > gpurun_out/pytest_g1.log 2>&1; echo "pytest train+config1 rc=$?"; tail -4 gpurun_out/pytest_g1.log
T="python tools/bench_train.py --regions 2 --cols 400 --solve --no-cublas"
SML_SOLVER=lu SML_LU_TIMING=1 timeout 300 $T > gpurun_out/bt_lu_g.log 2> gpurun_out/bt_lu_g.err; echo "bt lu dmma rc=$?"; cut -c1-400 gpurun_out/bt_lu_g.log; grep "sml lu" gpurun_out/bt_lu_g.err | head -4
SML_SOLVER=lu SML_LU_GEMM=fma SML_LU_TIMING=1 timeout 300 $T > gpurun_out/bt_lu_g2.log 2> gpurun_out/bt_lu_g2.err; echo "bt lu fma rc=$?"; grep "sml lu" gpurun_out/bt_lu_g2.err | head -4
SML_SOLVER=lu timeout 300 $T > gpurun_out/bt_lu_g3.log 2>> gpurun_out/bt_lu_g.err; echo "bt lu (no timing syncs) rc=$?"; cut -c1-400 gpurun_out/bt_lu_g3.log
SML_SOLVER=lu timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_lu_r01d.csv $T > gpurun_out/ncu_lu_g.log 2>&1; echo "ncu lu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_lu_r01d.csv", errors="ignore")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki].split("(")[0]; agg[k][0] += 1; agg[k][1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{k:40s} {c:6d} launches {t/1e6:10.3f} ms")
PY
