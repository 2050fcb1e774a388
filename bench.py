#!/usr/bin/env python
"""Headline benchmark: utterances/sec for batched greedy ASR (10 s synthetic log-mel, L = 128 decode steps).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (conv front-end -> encoder -> cross-K/V -> 128 greedy decode steps) over one
batch of synthetic utterances.  Workload = BASELINE.json configs[1] (C2): repo-default Speech-Transformer, batch 64
per GPU, fp16 hi|lo split tensor-core operands with fp32 accumulation.  Weak scaling: every GPU decodes its own 64 utterances,
no collective on the compute path, one final all_gather of the token matrices.

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events) with inputs resident in HBM; `e2e` goes through
the public API with pinned HOST buffers, H2D of the spectrogram and D2H of the transcripts inside the timed region.
`--impl reference` times the UNMODIFIED reference (`Transformer.evaluate`, imported from the git-ignored
baseline/_ref/ that tools/install_ref.sh fills; kind "reference") on the host cores, one utterance per step; only when
baseline/_ref is absent does it fall back to the oracle's literal restatement of model.py:125-151 (kind "port").

Beside the headline (C2, weak scaling) the same run reports four short extra passes under their own keys: `c3_strong`
(BASELINE config 3: paper-size model, global batch 256 split over the N GPUs), `c5_masks` (config 5: widened model,
mixed-length batch with key-padding masks on, length-balanced over the ranks), `c4_longform` (config 4: 30 s
utterances, 384 greedy steps, 64 utterances per GPU) and `cross_n_tokens` (every N decodes the same seed-1 global
batch; SHA-256 of the gathered tokens against the N=1 value in tests/golden/).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "C2"
METRIC = "utt/sec (10 s audio, greedy decode)"
UNIT = "utt/s"
DEC_CLASSES = ["dec_linear_qkv", "dec_attn_self", "dec_linear_out_proj", "dec_linear_cross_q", "dec_attn_cross",
               "dec_linear_ffn1", "dec_linear_ffn2", "dec_linear_classifier", "dec_select_embed"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: the workload's batch)")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel / per-phase profiling pass")
    ap.add_argument("--decode-mode", default="", choices=["", "cluster", "graph", "eager"],
                    help="sets ASR_B200_DECODE (default: the library default, cluster)")
    ap.add_argument("--no-extras", action="store_true", help="skip the c3_strong / c5_masks / c4_longform / cross_n_tokens passes")
    return ap.parse_args()


def coalesce_of(batch):
    """Batches per decode launch chosen by Transformer.greedy_decode_batches (coalesce=None)."""
    return max(1, min(4, 256 // max(1, batch)))


def workload_desc(cfg, batch, n_gpus):
    return {
        "workload": f"{cfg.name}: repo-default Speech-Transformer ({cfg.encoder_num_layers} enc / "
                    f"{cfg.decoder_num_layers} dec, d_model {cfg.embedding_dim}, {cfg.num_heads} heads, FFN {cfg.ff_dim}, "
                    f"vocab {cfg.vocab_size}), random init (seed 0, bf16-representable), synthetic structured log-mel "
                    f"{cfg.input_dim}x{cfg.frames} (10 ms hop), greedy decode exactly {cfg.decoder_seq_len} steps",
        "batch_per_gpu": batch, "global_batch": batch * n_gpus, "frames": cfg.frames,
        "encoder_frames": cfg.encoder_seq_len, "decode_steps": cfg.decoder_seq_len,
        "l2": "inputs rotate over 8 distinct device-resident batches (164 MB > 126 MB L2), no flush: consecutive steps "
              "overlap (encoder of the next steps under the decoder of the current ones); the serial pass in phase_ms "
              "flushes L2",
        "coalesce": "the serving loop (Transformer.greedy_decode_batches) decodes %d consecutive steps' batches per "
                    "launch (%d utterances, up to 8 per CTA cluster); every step's batch is fully processed and returned "
                    "separately" % (coalesce_of(batch), coalesce_of(batch) * batch),
        "parallelism": f"dp{n_gpus}: utterance sharding, one process per GPU, no collective on the compute path, "
                       "every rank downloads its own transcripts, ONE all_gather of the token ids at the end",
    }


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe), through NVML in this
    process: an `nvidia-smi -lms` loop beside the bench stalls the multi-stream launch path for milliseconds per
    sample (measured: -6 % throughput at 20 ms sampling), the two NVML calls used here do not."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index, period_s=0.02):
        self.idx, self.period, self.samples, self.reasons = gpu_index, period_s, [], set()
        self.stop_flag, self.thread, self.h, self.max_mhz = threading.Event(), None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            if "CUDA_VISIBLE_DEVICES" in os.environ and os.environ["CUDA_VISIBLE_DEVICES"].strip():
                vis = os.environ["CUDA_VISIBLE_DEVICES"].split(",")[self.idx].strip()
                self.h = pynvml.nvmlDeviceGetHandleByIndex(int(vis)) if vis.isdigit() else \
                    pynvml.nvmlDeviceGetHandleByUUID(vis)
            else:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.h = None
            return

        def loop():
            while not self.stop_flag.is_set():
                try:
                    mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                    rs = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    self.samples.append((time.perf_counter(), float(mhz), int(rs)))
                except Exception:
                    pass
                self.stop_flag.wait(self.period)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def clear(self):
        self.samples.clear()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        sm = [m for _, m, _ in self.samples]
        reasons = sorted({n for _, _, r in self.samples for n, bit in self.BAD.items() if r & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": reasons, "how": "NVML clock + clocks-event reasons every %d ms during the timed region"
                                           % int(self.period * 1e3)}


def file_sha(path):
    import hashlib
    return hashlib.sha256(open(path, "rb").read()).hexdigest()[:16] if os.path.exists(path) else None


def measured_traffic(kernel, utterances):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json:
    {kernel: {"dram_bytes", "utterances", "source", "kernel_source_sha"}}).  None when no capture of this kernel at this
    launch size has been committed OR the kernel's source changed since the capture (stale numbers are not reported)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        d = json.load(open(p)).get(kernel)
        if d and d.get("utterances", 64) == utterances:
            sha = file_sha(os.path.join(ROOT, "asr_transformer_b200", "csrc", "decode_cluster.cu"))
            if d.get("kernel_source_sha") in (None, sha):
                return d.get("dram_bytes")
    return None


def attention_tensor_pipe():
    """sm__pipe_tensor_cycles_active of the encoder flash-attention kernel from the newest committed encoder profile."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_encoder_kernels.md")))
    if not files:
        return None
    p = files[-1]
    rows = [r for r in (l.split("|") for l in open(p) if "attn_ts_kernel" in l or "attn_tc_kernel" in l)
            if len(r) >= 12]   # per-launch rows only
    if not rows:
        return None
    try:
        pct = [float(r[5]) for r in rows]
        us = [float(r[4]) for r in rows]
    except (ValueError, IndexError):
        return None
    return {"kernel": "attn_ts_kernel (encoder self attention)" if any("attn_ts" in r[1] for r in rows)
            else "attn_tc_kernel (encoder self attention)",
            "tensor_pipe_active_pct": round(sum(pct) / len(pct), 1), "us_per_launch": round(sum(us) / len(us), 1),
            "source": os.path.relpath(p, ROOT) + " (ncu --set full, cold cache; quoted, not measured in this run)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def decode_class_bytes(cfg, B):
    """Algorithmic HBM bytes per launch of each decode-step kernel class (fp16 weights and K/V caches, fp32
    activations), averaged over the L steps (self-attention reads t+1 cache rows at step t)."""
    D, FF, V, Tp, L = cfg.embedding_dim, cfg.ff_dim, cfg.vocab_size, cfg.encoder_seq_len, cfg.decoder_seq_len
    f = 4
    return {
        "dec_linear_qkv": 3 * D * D * 2 + B * D * f + B * 3 * D * f + B * 2 * D * 2,
        "dec_attn_self": B * (D * f + 2 * ((L + 1) / 2) * D * 2 + D * f),
        "dec_linear_out_proj": D * D * 2 + 3 * B * D * f,
        "dec_linear_cross_q": D * D * 2 + 2 * B * D * f,
        "dec_attn_cross": B * (D * f + 2 * Tp * D * 2 + D * f),
        "dec_linear_ffn1": D * FF * 2 + B * D * f + B * FF * f,
        "dec_linear_ffn2": D * FF * 2 + B * FF * f + 2 * B * D * f,
        "dec_linear_classifier": V * D * 2 + B * D * f + B * V * f,
        "dec_select_embed": B * V * f + B * D * f,
    }


def decode_alg_bytes(cfg, B):
    """algorithmic HBM bytes of ONE greedy decode of B utterances (all L steps): the 16-bit
    weights once per step, shared by the batch, + per utterance and step the cross K/V, the self K/V read (t + 1 rows at
    step t) and write, and the fp32 logits.  (11.138 MB + B x 1.933 MB per step at C2.)"""
    D, FF, V, Tp, L, nd = (cfg.embedding_dim, cfg.ff_dim, cfg.vocab_size, cfg.encoder_seq_len, cfg.decoder_seq_len,
                           cfg.decoder_num_layers)
    weights = 2 * (nd * (6 * D * D + 2 * D * FF) + D * V)
    per_utt = nd * 2 * Tp * D * 2 + nd * 2 * ((L + 1) / 2) * D * 2 + nd * 2 * D * 2 + 4 * V
    return {"weights_per_step": weights, "per_utterance_step": per_utt, "per_step": weights + B * per_utt,
            "per_decode": (weights + B * per_utt) * L}


def load_reference():
    """The unmodified reference Transformer class from baseline/_ref (tools/install_ref.sh), or None."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_root, "modules", "Transformer", "model.py")):
        return None
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from modules.Transformer.model import Transformer as RefTransformer
    return RefTransformer


def reference_evaluator(cfg, state):
    """Returns (fn(spectrum (b,1,F,T)) -> None running the reference's greedy path, kind)."""
    Ref = load_reference()
    if Ref is not None:
        torch.manual_seed(0)
        ref = Ref(**cfg.ctor_kwargs())
        ref.load_state_dict(state, strict=True)
        ref.eval()

        def fn(spec):
            with torch.no_grad():
                for b in range(spec.shape[0]):    # per utterance: evaluate() returns the last sample's tokens only (Q4)
                    ref.evaluate(spec[b:b + 1], torch.full((1, 1), cfg.bos_token_id, dtype=torch.int32))
        return fn, "reference"
    from oracle import speech_transformer as O

    def fn(spec):
        with torch.no_grad():
            O.evaluate_reference_style(state, spec, cfg)
    return fn, "port"


def run_reference(args, cfg):
    """Reference arm: the reference's own CPU implementation of the path (baseline/_ref: the unmodified
    Transformer.evaluate; the oracle's literal restatement only when that is absent) on the host cores.  One step = ONE
    utterance of the same workload (bounded sample; the reference costs O(L^2) per utterance)."""
    from asr_transformer_b200.workloads import build_model, cpu_state, structured_spectrum
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_model(cfg)
    sd = cpu_state(model)
    fn, kind = reference_evaluator(cfg, sd)
    batch = args.batch or cfg.batch
    spec = structured_spectrum(batch, cfg.frames, cfg.input_dim, seed=1)
    times = []
    for i in range(args.warmup + args.steps):
        b = i % batch
        t0 = time.perf_counter()
        fn(spec[b:b + 1])
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = len(times) / total
    sample = f"1 utterance per step ({cfg.frames} frames, {cfg.decoder_seq_len} decode steps), {len(times)} timed steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_desc(cfg, batch, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.decode_mode:
        os.environ["ASR_B200_DECODE"] = args.decode_mode
    mode = os.environ.get("ASR_B200_DECODE", "") or "cluster"
    from asr_transformer_b200 import workloads as W   # workload registry + synthetic inputs (no oracle on this arm)
    cfg = W.CONFIGS[args.workload]
    if args.impl == "reference":
        return run_reference(args, cfg)

    from asr_transformer_b200.workloads import build_model, cpu_state
    from asr_transformer_b200 import lib as L
    from asr_transformer_b200.parallel import gather_tokens

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    batch = args.batch or cfg.batch
    model = build_model(cfg, dev)
    lib = L.load()
    # every rank decodes its own utterances (different seed per rank): weak scaling
    spec_host = W.structured_spectrum(batch, cfg.frames, cfg.input_dim, seed=1 + rank).pin_memory()
    spec_dev = spec_host.to(dev)
    tokens = torch.empty(batch, cfg.decoder_seq_len + 1, dtype=torch.int32, device=dev)
    n_tok = torch.empty(batch, dtype=torch.int32, device=dev)
    enc = torch.empty(batch, cfg.encoder_seq_len, cfg.embedding_dim, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    eng = model._eng()

    mid = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def step_device(mark=None):
        eng.encode(spec_dev, out=enc)
        if mark is not None:
            mark.record()
        eng.decode_greedy(enc, tokens_out=tokens, n_tokens_out=n_tok)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ------------------------------------------------------------------ serial pass: one batch at a time, L2 flushed
    # between steps; gives the per-phase split and the non-overlapped step time
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for (s, e), m in zip(ev, mid):
        flush.fill_(1)                      # evict L2 (126 MB) between timed iterations
        s.record()
        step_device(m)
        e.record()
    barrier()
    ms = [s.elapsed_time(e) for s, e in ev]
    enc_ms = sum(s.elapsed_time(m) for (s, _), m in zip(ev, mid)) / args.steps
    dec_ms = sum(m.elapsed_time(e) for (_, e), m in zip(ev, mid)) / args.steps
    serial_ms = sum(ms) / args.steps

    # ------------------------------------------------------------------ device-timed throughput (inputs in HBM)
    # K steps through the pipelined serving loop: the encoder of step i+1 runs on the SMs the (latency-bound) cluster
    # decoder of step i leaves idle.  Inputs rotate over N_ROT distinct device-resident batches (> L2) instead of an L2
    # flush, which would serialise the steps.
    N_ROT = 8
    spec_rot = [spec_dev] + [W.structured_spectrum(batch, cfg.frames, cfg.input_dim, seed=1000 * (rank + 1) + i).to(dev)
                             for i in range(1, N_ROT)]

    def run_device(n):
        out = None
        for out in model.greedy_decode_batches((spec_rot[i % N_ROT] for i in range(n)), to_host=False):
            pass
        return out

    # the clock sampler is started BEFORE the warm-up: NVML start-up contends for the driver lock with the multi-stream
    # enqueue and would otherwise starve the first timed steps
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    sampler.start()
    time.sleep(0.5)
    run_device(max(args.warmup, 8))          # (two full groups untimed: the loop's buffers and allocator pools settle)
    barrier()
    sampler.clear()                         # keep only the samples taken during the timed region
    t_s, t_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.asr_launch_count()
    barrier()
    torch.cuda.cudart().cudaProfilerStart()     # `ncu --profile-from-start off` lists exactly the timed region
    t_s.record()
    run_device(args.steps)
    t_e.record()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop()
    launches = lib.asr_launch_count() - launches0
    total_ms = torch.tensor([t_s.elapsed_time(t_e)], dtype=torch.float64, device=dev)
    per_rank = None
    if dist is not None:
        # every rank's own timed region and SM clock: the headline divides by the MAX, this shows what the max is made of
        mine = torch.tensor([float(total_ms.item()), float(clocks.get("sm_mhz") or 0.0)], dtype=torch.float64, device=dev)
        allr = torch.empty(world * 2, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, 2).cpu()
        per_rank = {"ms": [round(float(v), 3) for v in allr[:, 0]], "sm_mhz": [float(v) for v in allr[:, 1]]}
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * batch * args.steps / (total_ms / 1e3)

    # ------------------------------------------------------------------ end to end through the public API
    # Every step uploads its own spectrogram batch from pinned host memory and downloads its transcripts; the public
    # serving call (Transformer.greedy_decode_batches) overlaps those copies with the neighbouring steps' compute.
    # Multi-GPU: every rank downloads ITS OWN transcripts per batch (plain D2H, no rendezvous in the loop);
    # the only collective is ONE all_gather of all the steps' token ids at the very end, inside the timed region.
    def run_e2e(n, trace=None):
        outs = []
        t_start = time.perf_counter()
        for out in model.greedy_decode_batches((spec_host for _ in range(n))):
            outs.append(out)
            if trace is not None:
                trace.append(round(1e3 * (time.perf_counter() - t_start), 2))
        if dist is not None and outs:
            tok = torch.cat([o[0] for o in outs], 0).to(dev, non_blocking=True)
            ntk = torch.cat([o[1] for o in outs], 0).to(dev, non_blocking=True)
            tok, ntk = gather_tokens(tok, ntk)
            run_e2e.gathered_rows = int(tok.shape[0])
            torch.cuda.synchronize()
        return outs[-1] if outs else None

    run_e2e(4)
    barrier()
    t0 = time.perf_counter()
    trace = []
    out_tokens, _ = run_e2e(args.steps, trace)
    barrier()
    if os.environ.get("ASR_B200_BENCH_TRACE"):
        print("e2e yield times (ms):", trace, file=sys.stderr)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * batch * args.steps / float(e2e_s.item())

    def step_serial():      # the same work without overlap: one blocking call per batch
        x = spec_host.to(dev, non_blocking=True)
        t, n = model.greedy_decode(x)
        return t.cpu(), n.cpu()

    step_serial()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_serial()
    barrier()
    ser_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ser_s, op=dist.ReduceOp.MAX)
    e2e_serial = world * batch * args.steps / float(ser_s.item())
    h2d = spec_host.numel() * 4
    d2h = out_tokens.numel() * 4 + out_tokens.shape[0] * 4

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 tensor-core operands (hi|lo split activations into every linear layer), fp32 accumulate",
        "data": "synthetic", "config": workload_desc(cfg, batch, world), "clocks": clocks, "per_rank": per_rank,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "Transformer.greedy_decode_batches (pinned host batches in, CPU transcripts out; copies of "
                       "neighbouring steps overlap compute; multi-GPU: per-rank D2H, one all_gather of the token ids at "
                       "the end of the run)", "serial_value": e2e_serial,
                "final_gather_rows": getattr(run_e2e, "gathered_rows", None)},
        "gpu_launches": int(launches),
        "phase_ms": {"conv_frontend+encoder": round(enc_ms, 3), "cross_kv+greedy_decode": round(dec_ms, 3),
                     "serial_step": round(serial_ms, 3),
                     "note": "one batch at a time with an L2 flush between steps; `value` overlaps consecutive steps"},
        "split_operands": int(lib.asr_split_operands()),
    }

    # ------------------------------------------------------------------ extra passes (all ranks): BASELINE configs 3 / 5
    # and the cross-N token identity check.  Short (2 timed steps each), device-timed, max over ranks.
    if not args.no_extras and args.workload == WORKLOAD:
        result.update(run_extras(model, dev, rank, world, dist, barrier))

    if rank == 0 and args.no_profile:
        print(json.dumps(result))
    elif rank == 0:
        # -------------------------------------------------------------- roofline of the dominant kernel
        # the decode launch of the timed region covers `co` coalesced batches: profile that launch
        co = coalesce_of(batch)
        pb = co * batch
        hbm_peak, tf_peak, peak_src = measured_peaks()
        prof_launches = {}
        for name_l, nb in (("bench_launch", pb), ("single_batch", batch)):
            enc_p = enc if nb == batch else torch.cat([eng.encode(spec_rot[i % N_ROT]) for i in range(co)], 0)
            tokens_p = torch.empty(nb, cfg.decoder_seq_len + 1, dtype=torch.int32, device=dev)
            ws = eng._ws(nb, 4 * cfg.encoder_seq_len + 3, cfg.decoder_seq_len)
            ms_cls = (C.c_float * 12)()
            n_cls = (C.c_int32 * 12)()
            phase = torch.zeros(max(148, torch.cuda.get_device_properties(dev).multi_processor_count), 16,
                                dtype=torch.int64, device=dev)
            for _ in range(2):     # second pass is the measured one (first warms caches / clocks)
                L.check(lib.asr_decode_profile(eng.handle, L.ptr(enc_p), nb, cfg.encoder_seq_len, cfg.decoder_seq_len,
                                               L.ptr(ws), ws.numel(), L.ptr(tokens_p), ms_cls, n_cls, L.ptr(phase),
                                               L.stream()), "asr_decode_profile")
            prof_launches[name_l] = (nb, list(ms_cls), list(n_cls), phase.double().cpu())
            if nb == batch and co == 1:
                prof_launches["bench_launch"] = prof_launches[name_l]
                break
        nb, ms_cls, n_cls, ph = prof_launches["bench_launch"]
        cp = ph[ph[:, 0] > 0]
        if len(cp):
            result["cluster_cycles_per_step"] = {k: round(float(cp[:, i].mean()) / cfg.decoder_seq_len, 1) for i, k in
                                                 enumerate(["total", "ring_wait", "exchange_wait", "producer_wait_empty",
                                                            "stages_total"])}
            result["cluster_ctas"] = int(len(cp))
            names_c = ["small_params+classifier+argmax", "ln1", "qkv_mm", "self_attention", "attn_finish(x2)",
                       "wo_mm(x2)", "all_reduce+ln(x3)", "cross_q_mm", "cross_attention", "ffn_w1_mm", "ffn_w2_mm"]
            result["cluster_phase_cycles_per_step"] = {n: round(float(cp[:, 5 + i].mean()) / cfg.decoder_seq_len, 1)
                                                       for i, n in enumerate(names_c)}
        # per-kernel classes of the graph (fallback) path, with that path's own per-kernel byte counts
        bytes_cls = decode_class_bytes(cfg, nb)
        prof = {}
        tot = sum(ms_cls[:9])
        for i, name in enumerate(DEC_CLASSES):
            if n_cls[i]:
                avg_us = 1e3 * ms_cls[i] / n_cls[i]
                prof[name] = {"launches": int(n_cls[i]), "avg_us": round(avg_us, 3), "share": round(ms_cls[i] / tot, 4),
                              "alg_bytes_per_launch": int(bytes_cls[name]),
                              "gbs": round(bytes_cls[name] / (avg_us * 1e-6) / 1e9, 1)}
        result["graph_path_kernel_profile"] = prof
        # the roofline of the dominant kernel: SURVEY.md 8(d) algorithmic bytes (weights once per step + per-utterance
        # K/V and logits) over the measured duration of ONE launch of dec_cluster_kernel
        alg = decode_alg_bytes(cfg, nb)
        result["decode_step"] = {"alg_bytes": int(alg["per_step"]), "weights_bytes": int(alg["weights_per_step"]),
                                 "per_utterance_bytes": int(alg["per_utterance_step"]), "utterances": nb,
                                 "roofline_ms_per_decode": round(alg["per_decode"] / (hbm_peak * 1e9) * 1e3, 3)}
        if mode[0] == "c" and n_cls[9]:
            gbs = alg["per_decode"] / (ms_cls[9] * 1e-3) / 1e9
            result["roofline"] = {"kernel": "dec_cluster_kernel (all %d decode steps of %d utterances, one launch)"
                                            % (cfg.decoder_seq_len, nb),
                                  "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                                  "frac": round(gbs / hbm_peak, 4),
                                  "traffic": measured_traffic("dec_cluster_kernel", nb) if args.workload == WORKLOAD else None,
                                  "peak_source": peak_src,
                                  "alg_bytes_per_launch": int(alg["per_decode"]), "ms_per_launch": round(ms_cls[9], 3),
                                  "how": "SURVEY.md 8(d) bytes per step x L / CUDA-event time of the launch on its stream"}
            if "single_batch" in prof_launches and prof_launches["single_batch"][0] != nb:
                nb1, ms1, n1, _ = prof_launches["single_batch"]
                alg1 = decode_alg_bytes(cfg, nb1)
                if n1[9]:
                    g1 = alg1["per_decode"] / (ms1[9] * 1e-3) / 1e9
                    result["roofline_single_batch"] = {"utterances": nb1, "ms_per_launch": round(ms1[9], 3),
                                                       "achieved": round(g1, 1), "frac": round(g1 / hbm_peak, 4),
                                                       "alg_bytes_per_launch": int(alg1["per_decode"])}
        else:
            top = max(prof, key=lambda k: prof[k]["share"])
            result["roofline"] = {"kernel": top, "bound": "hbm", "achieved": prof[top]["gbs"], "peak": hbm_peak,
                                  "unit": "GB/s", "frac": round(prof[top]["gbs"] / hbm_peak, 4), "traffic": None,
                                  "peak_source": peak_src}
        result["decode_mode"] = mode
        # second half of BASELINE.json's metric ("attention tensor-pipe %"): not measurable without a profiler, so it is
        # quoted from the committed ncu capture of this code, with its source
        att = attention_tensor_pipe()
        if att:
            result["attention_tensor_pipe"] = att

        # -------------------------------------------------------------- CPU baseline (the reference itself, host cores)
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            sd = cpu_state(model)
            fn, kind = reference_evaluator(cfg, sd)     # baseline/_ref (unmodified reference) or the oracle port
            n_cpu = 16
            fn(spec_host[:1].contiguous())              # warm-up: one utterance
            t0 = time.perf_counter()
            fn(spec_host[:n_cpu].contiguous())
            dt = time.perf_counter() - t0
            result["cpu_baseline"] = {
                "value": n_cpu / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                "sample": f"{n_cpu} utterances of the same workload through "
                          + ("the unmodified reference Transformer.evaluate (baseline/_ref)" if kind == "reference" else
                             "the oracle's literal restatement of the reference greedy loop")
                          + f" (model.py:125-151: per utterance, no KV cache), {dt:.1f} s"}
        print(json.dumps(result))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(model, dev, rank, world, dist, barrier):
    """BASELINE configs 3 and 5 and the cross-N token identity check, inside the same bench run (the driver only passes
    --gpus / --steps / --warmup).  Every pass is device-timed between barriers, max over ranks."""
    import hashlib
    from asr_transformer_b200 import workloads as W
    from asr_transformer_b200.parallel import (balanced_assignment, bucket_by_length, decode_sharded, gather_tokens,
                                               shard_range)
    out = {}

    def timed(fn, steps=2, warm=1):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    # ---- cross-N token identity: every N decodes the SAME seed-1 global batch of 64 utterances,
    # sharded contiguously over the ranks, gathered; the SHA-256 must equal the committed N=1 value
    c2 = W.CONFIGS["C2"]
    spec_g = W.structured_spectrum(64, c2.frames, c2.input_dim, seed=1)
    tok, ntk = decode_sharded(model, spec_g, dev)
    torch.cuda.synchronize()
    sha = hashlib.sha256(tok.cpu().to(torch.int32).contiguous().numpy().tobytes()).hexdigest()
    gp = os.path.join(ROOT, "tests", "golden", "cross_n_tokens.json")
    gold = json.load(open(gp)).get("C2_seed1_B64_sha256") if os.path.exists(gp) else None
    out["cross_n_tokens"] = {"workload": "C2, global batch 64 (seed 1), contiguous shards over %d rank(s)" % world,
                             "sha256": sha, "golden_n1": gold, "matches_golden": (sha == gold) if gold else None,
                             "distinct_rows": len({tuple(r) for r in tok.cpu().tolist()})}

    # ---- C3 strong scaling: paper-size model (12 encoder layers), global batch 256 split over the N GPUs
    c3 = W.CONFIGS["C3"]
    m3 = W.build_model(c3, dev)
    lo, hi = shard_range(c3.batch, rank, world)
    counts = [shard_range(c3.batch, r, world)[1] - shard_range(c3.batch, r, world)[0] for r in range(world)]
    x3 = W.structured_spectrum(c3.batch, c3.frames, c3.input_dim, seed=3)[lo:hi].to(dev)

    def step3():
        t, n = m3.greedy_decode(x3)
        if dist is not None:
            gather_tokens(t, n, counts)
    ms3 = timed(step3)
    out["c3_strong"] = {"workload": "C3: 12 enc / 6 dec, d_model 256, global batch 256 split over %d GPU(s) (%d per "
                                    "GPU), greedy decode 128 steps, inputs resident, token all_gather included" %
                                    (world, hi - lo),
                        "value": c3.batch / (ms3 / 1e3), "unit": UNIT, "ms_per_step": round(ms3, 3), "scaling": "strong"}
    del m3, x3
    torch.cuda.empty_cache()

    # ---- C5: widened model, mixed-length batch (len ~ U{400..1000}) with key-padding masks ON, 64 utterances per GPU,
    # length-balanced over the ranks, ONE padded batch per GPU.  Bucket size = the trade between encoder padding and decode
    # launches (tools/prof_c5.py, one B200, 64 utterances): a decode launch costs 12-15 ms whatever it holds, the whole
    # encode side 5-6 ms.  Buckets of 16 -> 4 launches, 53.3 ms; 32 -> 2 launches, 34.7 ms; 64 -> 1 launch, 27.7 ms (8
    # utterances per cluster of 8 CTAs; before that instance existed 64 utterances took two waves of clusters, 34.3 ms).
    c5 = W.CONFIGS["C5"]
    m5 = W.build_model(c5, dev)
    G = 64 * world
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(400, 1001, (G,), generator=g)
    mine = balanced_assignment(lens.tolist(), world)[rank]
    items = []
    for b in bucket_by_length([int(lens[i]) for i in mine], 64):
        idx = [mine[j] for j in b]
        ln = lens[idx]
        T = int(ln.max())
        x = W.structured_spectrum(len(idx), T, c5.input_dim, seed=500 + idx[0], lengths=ln).to(dev)
        items.append((x, ln.to(dev)))

    def step5():
        for x, ln in items:
            m5.greedy_decode(x, lengths=ln)
    ms5 = timed(step5)
    out["c5_masks"] = {"workload": "C5: 12 enc / 6 dec, d_model 512, 8 heads, FFN 2048; %d utterances per GPU, lengths "
                                   "U{400..1000} frames, key-padding masks on (encoder self attention + decoder cross "
                                   "attention), balanced_assignment over %d rank(s), one padded batch per GPU (one decode launch)" % (64, world),
                       "value": G / (ms5 / 1e3), "unit": UNIT, "ms_per_step": round(ms5, 3), "scaling": "weak",
                       "frames_processed_frac": round(float(sum(int(x.shape[-1]) * x.shape[0] for x, _ in items)) /
                                                      (64 * 1000), 3)}
    del m5, items
    torch.cuda.empty_cache()

    # ---- C4: long-form (30 s utterances -> 749 encoder frames, 384 greedy steps), 64 utterances per GPU.  The encoder's
    # self attention sees 749 keys here (attn_tc_kernel, online softmax), the cluster decoder 749 cross keys per utterance.
    c4 = W.CONFIGS["C4"]
    m4 = W.build_model(c4, dev)
    x4 = W.structured_spectrum(c4.batch, c4.frames, c4.input_dim, seed=40 + rank).to(dev)

    def step4():
        m4.greedy_decode(x4)

    def step4_loop():                       # the serving loop: 4 batches -> one decode launch of 256 utterances
        for _ in m4.greedy_decode_batches((x4 for _ in range(4)), to_host=False):
            pass
    ms4 = timed(step4, steps=1)
    ms4l = timed(step4_loop, steps=1)
    out["c4_longform"] = {"workload": "C4: 12 enc / 6 dec, d_model 256, batches of %d utterances of 30 s (T' = %d), greedy "
                                      "decode %d steps, inputs resident; `value`: 4 batches per GPU through the serving "
                                      "loop (one decode launch of 256 utterances), `single_batch_value`: one batch, "
                                      "Transformer.greedy_decode" % (c4.batch, c4.encoder_seq_len, c4.decoder_seq_len),
                          "value": 4 * c4.batch * world / (ms4l / 1e3), "unit": UNIT, "ms_per_step": round(ms4l / 4, 3),
                          "single_batch_value": c4.batch * world / (ms4 / 1e3), "single_batch_ms": round(ms4, 3),
                          "scaling": "weak"}
    del m4, x4
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
