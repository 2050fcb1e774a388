#!/bin/bash
# Encoder-side kernel table (one row per launch: time, tensor pipe, DRAM bytes, L2, occupancy) at 64, 256 and 512 utterances.
# usage: tools/run_profile_encoder.sh <tag>   -> gpurun_out/enc_kernels_<tag>_<B>.csv (small: raw metric CSV)
tag=${1:-r02}
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,sm__inst_executed_pipe_tensor.sum
for B in 64 256 512; do
  python tools/prof_encode.py --batch $B --reps 2 > gpurun_out/enc_plain_${tag}_$B.log 2>&1 && \
  ncu --metrics $M --clock-control none -k regex:'gemm|attn_t|conv|layernorm|f32_to|ffn' -s 26 -c 26 --csv \
      --log-file gpurun_out/enc_kernels_${tag}_$B.csv python tools/prof_encode.py --batch $B --reps 2 > gpurun_out/ncu_enc_${tag}_$B.log 2>&1
  echo "B=$B ncu exit $?"; tail -n 1 gpurun_out/enc_plain_${tag}_$B.log
done
