#!/bin/bash
# Quick A/B of the cluster decoder on a GPU box: token hash + decode time at 64/128/256 utterances, per-phase clocks at 256.
# usage: tools/quick_dec.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
{
python - <<'PY'
import hashlib, os, sys, torch
sys.path.insert(0, os.getcwd())
from asr_transformer_b200 import workloads as O
cfg = O.CONFIGS["C2"]
dev = torch.device("cuda", 0)
m = O.build_model(cfg, dev)
eng = m._eng()
for B in (64, 128, 256):
    spec = O.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=1).to(dev)
    enc = eng.encode(spec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(4):
        e0.record(); tok, n, _ = eng.decode_greedy(enc, max_len=128); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    h = hashlib.sha256(tok.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"B={B}: decode {min(ts[1:]):.3f} ms  tokens sha {h}")
PY
python tools/prof_phases.py 256 cluster
} > gpurun_out/quick_$tag.log 2>&1
cat gpurun_out/quick_$tag.log
