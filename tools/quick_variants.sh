#!/bin/bash
# usage: tools/quick_variants.sh name1 name2 ...   (libs from tools/variants.py; "stock" = the in-tree library)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = stock ]; then unset ASR_B200_LIB; else export ASR_B200_LIB=$PWD/asr_transformer_b200/build/variants/libasr_$v.so; fi
  echo "=== $v"
  bash tools/quick_dec.sh $v | grep -v "^cluster B="
done
