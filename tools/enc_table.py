#!/usr/bin/env python
"""gpurun_out/enc_kernels_<tag>_<B>.csv (tools/run_profile_encoder.sh: ncu --metrics ... --csv) -> markdown table, one row
per launch.  python tools/enc_table.py csv [csv ...] > profiles/<name>.md"""
import csv
import io
import re
import sys
from collections import OrderedDict


def rows_of(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    out = OrderedDict()
    for r in rd:
        k = r["ID"]
        d = out.setdefault(k, {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        if name.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[name] = v
    return out


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"asr::<unnamed>::", "", n)
    return re.sub(r"\(.*$", "", n)[:60]


for path in sys.argv[1:]:
    print(f"## {path}\n")
    print("| kernel | grid | block | time us | tensor pipe active % | DRAM GB (r+w) | DRAM GB/s | L2 % | warps active % | regs | CTAs/SM (smem) |")
    print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    tot = 0.0
    fam = OrderedDict()
    for d in rows_of(path).values():
        t = d.get("gpu__time_duration.sum", 0.0)
        by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        tot += t
        f = re.sub(r"<.*", "", short(d["name"]))
        a = fam.setdefault(f, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        print(f"| `{short(d['name'])}` | {d['grid']} | {d['block']} | {t:.1f} | "
              f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0):.1f} | {by / 1e9:.3f} | "
              f"{by / 1e9 / (t * 1e-6) if t else 0:.0f} | {d.get('lts__throughput.avg.pct_of_peak_sustained_elapsed', 0.0):.1f} | "
              f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0.0):.1f} | "
              f"{int(d.get('launch__registers_per_thread', 0))} | {int(d.get('launch__occupancy_limit_shared_mem', 0))} |")
    print(f"\ntotal {tot:.1f} us over the captured launches\n")
    print("| kernel family | launches | total us | share | time-weighted tensor pipe % |")
    print("|---|---:|---:|---:|---:|")
    for f, (n, t, w) in fam.items():
        print(f"| `{f}` | {n} | {t:.1f} | {t / tot:.3f} | {w / t if t else 0:.1f} |")
    print()
