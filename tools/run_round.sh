#!/bin/bash
# Round check on a GPU box: all GPU tests, the default bench, the profile captures (tools/run_profile.sh <tag>)
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_$tag.log
timeout 400 python bench.py > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "bench exit $?"
bash tools/run_profile.sh $tag
