mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_r02a.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_r02a.log
timeout 400 python bench.py > gpurun_out/bench_r02a.log 2> gpurun_out/bench_r02a.err; echo "bench exit $?"
bash tools/run_profile.sh r02a
