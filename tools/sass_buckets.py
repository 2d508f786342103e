#!/usr/bin/env python
"""Executed thread instructions per element by PHASE of move_kernel (buckets of source lines), from an ncu source-page CSV joined with
nvdisasm line info.  usage: sass_buckets.py <ncu_source.csv> <cubin> <mangled function> <elements>"""
import collections, csv, re, subprocess, sys
csvp, cubin, fun, nelem = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = open(__file__.replace("tools/sass_buckets.py", "pyfilter_b200/csrc/move.cuh")).read().splitlines()
def find(marker):
    for i, l in enumerate(src):
        if marker in l:
            return i + 1
    raise KeyError(marker)
L = {k: find(v) for k, v in dict(weights="double mv_weights(", mark="void mv_mark(", remark="void mv_remark(", emit="int32_t mv_emit(",
                                 kernel="move_kernel(MoveArgs c)", loads="the tile comes on chip", resample="if (resampled) {",
                                 window="auto run_window", after="if (n_out > n_in) {", partial="per-tile partial record").items()}
def bucket(f, ln):
    if f == "move.cuh":
        if ln < L["mark"]: return "1 weights"
        if ln < L["remark"]: return "3 count+mark"
        if ln < L["emit"]: return "9 remark"
        if ln < L["kernel"]: return "4 emit"
        if ln < L["loads"]: return "0 prologue"
        if ln < L["resample"]: return "0 loads"
        if ln < L["window"]: return "2 scan/publish/look-back"
        if ln < L["after"]: return "5 propagate"
        if ln < L["partial"]: return "4 emit"
        return "7 partial+finalize"
    if f == "common.cuh": return "1 weights" if 60 <= ln <= 100 else "2 scan/publish/look-back"
    if f == "exact_scan.h": return "3 count+mark"
    if f in ("philox.h",): return "5a philox+normal"
    if f == "models.h": return "5b model densities"
    if f == "step.cuh":
        if ln < 200: return "7 partial+finalize"
        if 200 <= ln < 300: return "5b model densities"
        if 576 <= ln < 660: return "5c soft-max statistics"
        if ln >= 420 and ln < 576: return "7 partial+finalize"
        return "5 propagate"
    if f == "resample.cuh": return "2 scan/publish/look-back"
    return "8 other (" + f + ")"
dis_all = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
start = dis_all.index(f".text.{fun}:")
end = dis_all.find("//--------------------- .text.", start)
lines, cur = [], ("?", 0)
for ln in dis_all[start:end if end > 0 else None].splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln):
        lines.append(cur)
rows = list(csv.reader(open(csvp)))
hdr = [r for r in rows if r and r[0] == "Address"][0]
data = [r for r in rows if r and r[0].startswith("0x")][:len(lines)]
ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, sagg = collections.Counter(), collections.Counter()
for l, d in zip(lines, data):
    b = bucket(*l)
    agg[b] += int(d[ie].replace(",", "") or 0); sagg[b] += int(d[sm].replace(",", "") or 0)
tot, stot = sum(agg.values()), sum(sagg.values())
print(f"thread instructions per element: {tot * 32 / nelem:.1f}")
for k in sorted(agg):
    print(f"{agg[k] * 32 / nelem:7.2f} instr/elt  {agg[k] / tot:6.1%}   stall samples {sagg[k] / max(stot, 1):6.1%}   {k}")
