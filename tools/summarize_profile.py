#!/usr/bin/env python
"""profiles/ summary from one gpurun capture: ncu launch list (gpu__time_duration) + `--set full` raw page.
usage: summarize_profile.py <launches.csv> <raw.csv> <tag>  -> markdown on stdout"""
import csv, collections, sys
launches, raw, tag = sys.argv[1:4]
rows = list(csv.reader(open(launches)))
i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[i0]; k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) <= v: continue
    name = r[k].split("(")[0].replace("void ", "")
    agg.setdefault(name, []).append(float(r[v].replace(",", "")) / 1e3)
tot = sum(sum(x) for x in agg.values())
print(f"# {tag} - ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised: compare shares)\n")
print("| kernel | launches | avg us | share of device time |\n|---|---|---|---|")
for n, x in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| {n} | {len(x)} | {sum(x)/len(x):.2f} | {100*sum(x)/tot:.1f}% |")
rr = list(csv.reader(open(raw)))
h = rr[0]
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_bytes.sum", "L2 bytes"), ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
print(f"\n# {tag} - `ncu --set full` per launch (one launch each; caches flushed between replays)\n")
print("| kernel | " + " | ".join(w[1] for w in want) + " |\n|" + "---|" * (len(want) + 1))
units = rr[1]
for r in rr[2:]:
    cells = []
    for m, _ in want:
        if m in h:
            i = h.index(m); val = r[i]; u = units[i]
            try: val = f"{float(val):.4g}"
            except ValueError: pass
            cells.append(f"{val} {u}".strip())
        else: cells.append("-")
    print("| " + r[h.index("Kernel Name")].split("(")[0].replace("void ", "") + " | " + " | ".join(cells) + " |")
