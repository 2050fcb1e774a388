mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","phase_ms","roofline","cluster_cycles_per_step","cluster_phase_cycles_per_step","decode_kernels_ms"]: print(k, d.get(k))
PY
tail -n 5 gpurun_out/bench.err
