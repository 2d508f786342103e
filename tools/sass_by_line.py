#!/usr/bin/env python
"""Join an ncu SASS source-page CSV (per-instruction executed counts / stall samples) with nvdisasm line info of the same
cubin and aggregate by CUDA source line.   usage: sass_by_line.py <ncu_sass.csv> <cubin> <mangled function> [top]"""
import collections, csv, re, subprocess, sys

def main():
    csv_path, cubin, fun = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-g", "-c", "-fun", fun, cubin], capture_output=True, text=True).stdout
    if not dis.strip():
        dis_all = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        start = dis_all.index(f".text.{fun}:")
        end = dis_all.find("//--------------------- .text.", start)
        dis = dis_all[start:end if end > 0 else None]
    lines, cur = [], ("?", 0)
    for ln in dis.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    rows = list(csv.reader(open(csv_path)))
    # first kernel instance only
    hdr, data, seen = None, [], 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            seen += 1
            if seen > 1:
                break
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
            data.append(dict(zip(hdr, r)))
    print(f"sass instructions: disasm {len(lines)}  ncu {len(data)}")
    n = min(len(lines), len(data))
    agg = collections.defaultdict(lambda: [0, 0])
    for i in range(n):
        d = data[i]
        ie = int(d["Instructions Executed"].replace(",", "") or 0)
        sm = int(d["# Samples"].replace(",", "") or 0)
        agg[lines[i]][0] += ie
        agg[lines[i]][1] += sm
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot}  stall samples {tots}")
    byfile = collections.defaultdict(lambda: [0, 0])
    for k, v in agg.items():
        byfile[k[0]][0] += v[0]; byfile[k[0]][1] += v[1]
    for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  file {k:24s} inst {v[0]/tot:6.1%} samples {v[1]/tots:6.1%}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"inst {v[0]/tot:6.1%}  samples {v[1]/tots:6.1%}  {k[0]}:{k[1]}")

main()
