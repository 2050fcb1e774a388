mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_r02k.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu_r02k.log
timeout 400 python bench.py > gpurun_out/bench_r02k.log 2> gpurun_out/bench_r02k.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02k.log").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","phase_ms","cross_n_tokens","clocks","gpu_launches","c3_strong","c5_masks"]: print(k, d.get(k))
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
