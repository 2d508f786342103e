#!/usr/bin/env python
"""Static SASS statistics of one kernel, no GPU needed: instructions per CUDA source line (nvdisasm line info; inlined code is
attributed to the innermost line), per opcode, and where the local-memory (spill) accesses sit.
usage: sass_static.py <lib.so | cubin> <substring of the mangled function name> [top] [lo-hi address range in hex]"""
import collections, os, re, subprocess, sys, tempfile


def cubin_of(path):
    if path.endswith(".cubin"):
        return path
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, check=True, capture_output=True)
    return os.path.join(d, sorted(os.listdir(d))[0])


def main():
    cubin = cubin_of(sys.argv[1])
    key = sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rng = None
    if len(sys.argv) > 4:
        lo, hi = sys.argv[4].split("-")
        rng = (int(lo, 16), int(hi, 16))
    dis_all = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    funs = re.findall(r"\.text\.(\S+):", dis_all)
    fun = [f for f in funs if key in f]
    if not fun:
        sys.exit(f"no function matching {key}; have {len(funs)} functions")
    fun = fun[0]
    start = dis_all.index(f".text.{fun}:")
    end = dis_all.find("//--------------------- .text.", start)
    dis = dis_all[start:end if end > 0 else None]
    cur = ("?", 0)
    by_line, by_op, spills = collections.Counter(), collections.Counter(), collections.Counter()
    total = 0
    for ln in dis.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if rng and not (rng[0] <= addr < rng[1]):
            continue
        t = m.group(2).split()
        op = t[1] if t[0].startswith("@") else t[0]
        total += 1
        by_line[cur] += 1
        by_op[op.split(".")[0]] += 1
        if op.startswith(("STL", "LDL")):
            spills[cur] += 1
    print(f"{fun}: {total} static instructions")
    for k, v in by_line.most_common(top):
        print(f"{v:6d}  {k[0]}:{k[1]}" + (f"   (local-memory accesses: {spills[k]})" if spills[k] else ""))
    print()
    print("  ".join(f"{k} {v}" for k, v in by_op.most_common(30)))
    if spills:
        print("local-memory accesses by line:", dict((f"{k[0]}:{k[1]}", v) for k, v in spills.most_common(20)))


main()
