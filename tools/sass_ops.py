#!/usr/bin/env python
"""Per-source-line and per-opcode executed-instruction counts of one kernel: joins an ncu source-page CSV with nvdisasm.
usage: sass_ops.py <ncu_source.csv> <cubin> <mangled function> <elements per launch>"""
import subprocess, re, csv, collections, sys
csvp, cubin, fun, nelem = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
dis_all = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
start = dis_all.index(f".text.{fun}:")
end = dis_all.find("//--------------------- .text.", start)
dis = dis_all[start:end if end > 0 else None]
lines, cur, ops = [], ("?", 0), []
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        lines.append(cur); ops.append(m.group(1))
rows = list(csv.reader(open(csvp)))
hdr = [r for r in rows if r and r[0] == "Address"][0]
data = [r for r in rows if r and r[0].startswith("0x")][:len(lines)]
ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, opagg, sagg, tot, stot = collections.Counter(), collections.Counter(), collections.Counter(), 0, 0
for l, o, d in zip(lines, ops, data):
    n = int(d[ie].replace(",", "") or 0); s = int(d[sm].replace(",", "") or 0)
    tot += n; stot += s; agg[l] += n; sagg[l] += s
    t = o.split()
    op = t[1] if t[0].startswith("@") else t[0]
    opagg[op.split(".")[0]] += n
print(f"thread instructions per element: {tot * 32 / nelem:.1f}   (warp instructions {tot}, samples {stot})")
for k, v in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 30):
    print(f"{v * 32 / nelem:7.2f} instr/elt  samples {sagg[k] / max(stot, 1):6.1%}  {k[0]}:{k[1]}")
print()
for k, v in opagg.most_common(25):
    print(f"{v * 32 / nelem:7.2f}  {k}")
