#!/bin/bash
# Round profile capture on a GPU box: launch list of the bench command + one full capture of the dominant kernel
# (each only after the plain command has exited 0).  usage: tools/run_profile.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-profile"
$CMD > gpurun_out/ncu_plain_$tag.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv \
  --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1; echo "ncu list exit $?"
PD="python tools/prof_decode.py --batch 256 --reps 1"
$PD > gpurun_out/prof_plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dec_cluster -c 1 \
  -o gpurun_out/prof_cluster_$tag $PD > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/
