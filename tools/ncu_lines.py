#!/usr/bin/env python
"""Attribute an ncu capture's per-instruction samples to CUDA source lines.

ncu's CLI only prints per-SASS-instruction metrics; this joins them with the line
table of the same kernel (nvdisasm -g on the object that was profiled) and prints
time share / stall mix per source line and per coarse region of the sweep.

usage: ncu_lines.py <report.ncu-rep> <object-or-cubin> <kernel-regex> [--top N]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_linemap(obj, kernel_rx):
    tmp = tempfile.mkdtemp()
    if not obj.endswith(".cubin"):
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True,
                       capture_output=True)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    else:
        cubins = [obj]
    rx = re.compile(kernel_rx)
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        cur_fn, cur_line, inl = None, None, None
        maps = {}
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                maps.setdefault(cur_fn, {})
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
            if m and cur_fn:
                maps[cur_fn][int(m.group(1), 16)] = (cur_line, m.group(2).strip())
        for fn, mp in maps.items():
            if rx.search(fn) and mp:
                return fn, mp
    raise SystemExit("kernel not found in " + obj)


def ncu_source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(rows[2][0], 16)
    data = []
    for r in rows[2:]:
        st = {h[6:]: int(r[ix[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        data.append((int(r[0], 16) - base, r[1].strip(), int(r[ix["# Samples"]]),
                     int(r[ix["Instructions Executed"]]), st))
    return data


def main():
    rep, obj, krx = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    fn, lm = sass_linemap(obj, krx)
    data = ncu_source(rep)
    by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = sum(d[2] for d in data)
    tot_i = sum(d[3] for d in data)
    miss = 0
    for off, sass, smp, nexe, st in data:
        ent = lm.get(off)
        if ent is None or ent[0] is None:
            miss += smp
            key = ("?", 0)
        else:
            key = ent[0]
        b = by_line[key]
        b[0] += smp
        b[1] += nexe
        b[2].update(st)
    print("kernel", fn[:90])
    print("samples %d  warp-instructions %d  (unmapped samples %d)" % (tot_s, tot_i, miss))
    print("%-28s %7s %6s %10s  top stalls" % ("file:line", "samples", "%", "instr"))
    for key, (smp, nexe, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join("%s %d" % (k, v) for k, v in st.most_common(3) if v)
        print("%-28s %7d %6.2f %10d  %s" % ("%s:%d" % key, smp, 100.0 * smp / max(tot_s, 1), nexe, tops))
    allst = collections.Counter()
    for _, _, _, _, st in data:
        allst.update(st)
    print("stall mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(sum(allst.values()), 1))
                                  for k, v in allst.most_common(8)))


if __name__ == "__main__":
    main()
