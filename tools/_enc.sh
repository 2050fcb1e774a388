timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for v in 1 0; do export ASR_B200_LN_TMA=$v; echo -n "LN_TMA=$v: "; ncu --metrics $M --clock-control none -k regex:'ffn_fused' -s 2 -c 3 python tools/prof_encode.py --batch 256 --reps 2 2>&1 | grep -E "gpu__time|tensor_cycles" | awk '{print $(NF)}' | paste - - - - - -; done
unset ASR_B200_LN_TMA
python tools/prof_encode.py --batch 256 --reps 4 | tail -1
python tools/prof_encode.py --batch 64 --reps 4 | tail -1
