mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 120 -k "attention or mha" > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_attn.log
timeout 120 python tools/prof_attn.py 2>&1 | tail -4
