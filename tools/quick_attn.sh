mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_r02i.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu_r02i.log
timeout 400 python bench.py > gpurun_out/bench_r02i.log 2> gpurun_out/bench_r02i.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02i.log").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","phase_ms","roofline","cross_n_tokens","clocks"]: print(k, d.get(k))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ts -s 3 -c 1 -f -o gpurun_out/prof_attn_ts2 python tools/prof_attn.py --batches 256 --reps 3 > gpurun_out/ncu_attn_ts2.log 2>&1; echo "ncu exit $?"
