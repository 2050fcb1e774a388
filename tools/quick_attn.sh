mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_r02j.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu_r02j.log
timeout 400 python bench.py --steps 40 --warmup 4 > gpurun_out/bench_r02j.log 2> gpurun_out/bench_r02j.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02j.log").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","phase_ms","cross_n_tokens","clocks","gpu_launches"]: print(k, d.get(k))
PY
bash tools/run_profile_encoder.sh r02j 2>&1 | tail -6
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-profile"
$CMD > gpurun_out/ncu_plain_r02j.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/launches_r02j.csv $CMD > gpurun_out/ncu_list_r02j.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ts -s 3 -c 1 -f -o gpurun_out/prof_attn_r02j python tools/prof_attn.py --batches 256 --reps 3 > gpurun_out/ncu_attn_r02j.log 2>&1; echo "ncu attn exit $?"
