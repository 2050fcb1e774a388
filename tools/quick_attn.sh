#!/bin/bash
# Quick check of the attention kernels on a GPU box: operator parity tests, then the op alone timed with CUDA events
# (C2 encoder shape, 64 / 256 / 512 utterances).   gpurun -- 'bash tools/quick_attn.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 120 -k "attention or mha" > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_attn.log
timeout 120 python tools/prof_attn.py 2>&1 | tail -3
