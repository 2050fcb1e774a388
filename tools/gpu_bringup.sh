#!/bin/bash
# Bring-up on a GPU box: each group runs in its own process so one faulting kernel cannot poison the rest.
# usage: tools/gpu_bringup.sh [groups...]   (default: all)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 1800 python -m pytest "$@" -q --tb=short -p no:cacheprovider --timeout 400 -s > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; }
groups=${@:-"probe ln gemm attn modules dec model smoke bench"}
for g in $groups; do
case $g in
  probe) run probe tests/test_ops_gpu.py -m gpu -k "umma_probe" ;;
  ln) run ln tests/test_ops_gpu.py -m gpu -k "layernorm or embed" ;;
  gemm) run gemm tests/test_ops_gpu.py -m gpu -k "gemm" ;;
  attn) run attn tests/test_ops_gpu.py -m gpu -k "attention and not dec_attention" ;;
  modules) run modules tests/test_ops_gpu.py -m gpu -k "module or conv" ;;
  dec) run dec tests/test_ops_gpu.py -m gpu -k "dec_" ;;
  model) run model tests/test_model_gpu.py -m gpu ;;
  smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log ;;
  bench) timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err ;;
  benchgraph) timeout 600 python bench.py --steps 5 --warmup 3 --decode-mode graph --no-cpu-baseline > gpurun_out/benchgraph.log 2> gpurun_out/benchgraph.err; echo "benchgraph exit $?"; tail -c 3000 gpurun_out/benchgraph.log; tail -n 5 gpurun_out/benchgraph.err ;;
  ncu) CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile"
       $CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
       $CMD > gpurun_out/ncu_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dec_persistent -c 1 -o gpurun_out/prof_persistent $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep ;;
  benchteams) for t in 1 2 8 16; do ASR_B200_TEAMS=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_t$t.log 2>&1; echo "teams=$t: $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_t$t.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])")"; done ;;
  benchstream) timeout 600 python bench.py --steps 5 --warmup 3 --decode-mode stream --no-cpu-baseline > gpurun_out/benchstream.log 2> gpurun_out/benchstream.err; echo "benchstream exit $?"; python -c "import json; d=json.loads(open('gpurun_out/benchstream.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['decode_kernels_ms'], d['roofline'])"; tail -n 3 gpurun_out/benchstream.err ;;
  cluster) run cluster tests/test_model_gpu.py -m gpu -k "cluster" ;;
  benchcluster) timeout 600 python bench.py --steps 5 --warmup 3 --decode-mode cluster --no-cpu-baseline > gpurun_out/benchcluster.log 2> gpurun_out/benchcluster.err; echo "benchcluster exit $?"; python -c "import json; d=json.loads(open('gpurun_out/benchcluster.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['decode_kernels_ms'], d['roofline'], d.get('cluster_cycles_per_step'), d.get('cluster_ctas'), d.get('cluster_phase_cycles_per_step'))"; tail -n 3 gpurun_out/benchcluster.err ;;
  benchref) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/benchref.log 2>&1; echo "benchref exit $?"; tail -c 1500 gpurun_out/benchref.log ;;
esac
done
