export ASR_B200_LIB=$PWD/asr_transformer_b200/build/variants/libasr_late.so
export ASR_B200_CLUSTER_GU=8
for B in 8 256; do python tools/prof_phases.py $B cluster 2>&1 | tail -3; done
