import os, sys, torch
sys.path.insert(0, os.getcwd())
from asr_transformer_b200 import workloads as W
cfg = W.CONFIGS["C2"]; dev = torch.device("cuda", 0)
m = W.build_model(cfg, dev)
specs = [W.structured_spectrum(64, cfg.frames, cfg.input_dim, seed=1 + i).to(dev) for i in range(8)]
def run(n, co):
    for out in m.greedy_decode_batches((specs[i % 8] for i in range(n)), to_host=False, coalesce=co): pass
for co in (2, 3, 4, 2, 4):
    run(12, co); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(48, co); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"coalesce {co}: {ms/48:.3f} ms/step = {48*64/ms*1e3:.0f} utt/s")
