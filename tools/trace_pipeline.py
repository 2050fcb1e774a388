import os, sys, torch
sys.path.insert(0, os.getcwd())
from asr_transformer_b200 import workloads as O  # noqa: E402  (workload registry + synthetic inputs)
from asr_transformer_b200.workloads import build_model
cfg = O.CONFIGS["C2"]; dev = torch.device("cuda", 0)
m = build_model(cfg, dev)
specs = [O.structured_spectrum(64, cfg.frames, cfg.input_dim, seed=1 + i).to(dev) for i in range(8)]
eng = m._eng()
# monkeypatch engine calls to record events
ev = []
orig_enc, orig_dec = eng.encode, eng.decode_greedy
def enc(x, *a, **k):
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    r = orig_enc(x, *a, **k)
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    ev.append(("enc", e0, e1)); return r
def dec(x, *a, **k):
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    r = orig_dec(x, *a, **k)
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    ev.append(("prep" if k.get("phase") == "prepare" else "dec", e0, e1)); return r
eng.encode, eng.decode_greedy = enc, dec
def run(n):
    for out in m.greedy_decode_batches((specs[i % 8] for i in range(n)), to_host=False): pass
run(6); torch.cuda.synchronize(); ev.clear()
t0 = torch.cuda.Event(enable_timing=True); t0.record()
run(12); torch.cuda.synchronize()
for name, a, b in ev:
    print(f"{name:5s} start {t0.elapsed_time(a):8.3f} end {t0.elapsed_time(b):8.3f} dur {a.elapsed_time(b):7.3f}")
