mkdir -p gpurun_out
timeout 120 python tools/prof_attn.py --batches 256 --reps 5 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 3 -c 1 -f -o gpurun_out/prof_attn_pp python tools/prof_attn.py --batches 256 --reps 3 > gpurun_out/ncu_attn_pp.log 2>&1; echo "ncu exit $?"
