#!/usr/bin/env python
"""profiles/<tag>_ncu_raw_{fwd,bwd}.csv (ncu --set full --page raw --csv) -> profiles/<tag>_ncu_summary.md."""
import csv, sys
tag = sys.argv[1]
cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd MB"), ("dram__bytes_write.sum", "wr MB"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"), ("launch__registers_per_thread", "regs")]
out = [f"# Round 1, capture {tag}: `ncu --set full --clock-control none` of the bench step\n",
       "Two windows of 24 consecutive hot-kernel launches of the last step of `python bench.py --steps 1 --warmup 3 --quick` "
       f"(`scripts/gpu_profile.sh {tag} full`): the level-0 forward (first KPConv, unary blocks, the 32-channel KPConv of the "
       "first bottleneck) and the level-0 end of the backward pass.  Full text pages: "
       f"`{tag}_ncu_details_fwd.txt`, `{tag}_ncu_details_bwd.txt`; raw metric pages: `{tag}_ncu_raw_*.csv`.  "
       "DRAM GB/s = (read + written bytes) / duration of that launch (cold cache, serialised).\n"]
for part in ("fwd", "bwd"):
    rows = list(csv.reader(open(f"profiles/{tag}_ncu_raw_{part}.csv")))
    hdr, units = rows[0], rows[1]
    idx = {n: hdr.index(n) for n, _ in cols if n in hdr}
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    out.append(f"\n## {part}\n")
    out.append("| kernel | grid | " + " | ".join(h for n, h in cols if n in idx) + " | DRAM GB/s |")
    out.append("|---|---|" + "---:|" * (len(idx) + 1))
    def to(v, u, target):
        v = float(v.replace(",", ""))
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
        return v * scale.get(u, 1.0)
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("mvk::<unnamed>::", "").replace("unnamed>::", "")[:40]
        vals, us, mb = [], 0.0, 0.0
        for n, h in cols:
            if n not in idx:
                continue
            v = to(r[idx[n]], units[idx[n]], h) if r[idx[n]] not in ("", "n/a") else float("nan")
            if h == "us":
                us = v
            if h in ("rd MB", "wr MB"):
                mb += v
            vals.append(f"{v:.1f}" if h not in ("regs",) else f"{int(v)}")
        out.append(f"| `{name}` | {r[gi]} | " + " | ".join(vals) + f" | {mb / us * 1e3:.0f} |")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
