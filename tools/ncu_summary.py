#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small, committable text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_prof.md   (needs `ncu` on PATH, no GPU)
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEY_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:90]


def launches(src, dst, title=""):
    rows = [l for l in open(src) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    total = 0.0
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n{title}\n\n")
        f.write("ncu `--metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and serialised; compare SHARES.\n\n")
        f.write("| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---:|---:|---:|---:|\n")
        for (k, g, b), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {g} | {b} | {n} | {ns/1e3:.1f} | {ns/1e3/n:.2f} | {ns/total:.3f} |\n")
        f.write(f"\ntotal {total/1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        fam = OrderedDict()
        for (k, g, b), (n, ns) in agg.items():
            name = re.sub(r"<.*$", "", k.replace("<unnamed>::", "")).split("::")[-1]
            fa = fam.setdefault(name, [0, 0.0])
            fa[0] += n
            fa[1] += ns
        f.write("\n| kernel family | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, ns) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns/1e3:.1f} | {ns/total:.3f} |\n")


def full(src, dst, title=""):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n{title}\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"## {short(d.get('Kernel Name',''))}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in KEY_METRICS:
                if m in d:
                    f.write(f"| {m} | {d[m]} | {u[m]} |\n")
            f.write("\n")


def table(src, dst, title=""):
    """One row per captured launch: duration, tensor-pipe activity, DRAM throughput, occupancy."""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full, one row per launch ({src})\n\n{title}\n\n")
        f.write("| kernel | grid | block | time us | tensor pipe active % | DRAM GB/s (read+write) | DRAM % of peak | L2 % | "
                "regs | warps active % | CTAs/SM limit (smem) |\n|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))

            def g(k, default=0.0):
                try:
                    return float(d.get(k, default))
                except ValueError:
                    return default
            t_us = g("gpu__time_duration.sum")
            unit = dict(zip(hdr, rows[1])).get("gpu__time_duration.sum", "us")
            t_us *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
            units = dict(zip(hdr, rows[1]))
            def to_bytes(k):
                return g(k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units.get(k, "byte"), 1)
            by = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            f.write(f"| `{short(d.get('Kernel Name', ''))}` | {d.get('Grid Size')} | {d.get('Block Size')} | {t_us:.1f} | "
                    f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                    f"{by / max(t_us, 1e-9) / 1e3:.0f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                    f"{g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {int(g('launch__registers_per_thread'))} | "
                    f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                    f"{int(g('launch__occupancy_limit_shared_mem'))} |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else ""
    {"launches": launches, "full": full, "table": table}[mode](src, dst, title)
