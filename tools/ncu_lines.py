#!/usr/bin/env python
"""Aggregate an ncu source-page export (SASS, per-instruction warp-stall samples) by CUDA source line.

ncu's CSV export of the source page lists SASS instructions without their source lines; nvdisasm -g on the same cubin does
carry them, in the same instruction order.  Usage:
  ncu -i rep.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g -c icp.sm_100a.cubin > all.dis
  python tools/ncu_lines.py sass.csv all.dis '<kernel substring in the csv>' '<mangled substring in the disassembly>' [top]
"""
import csv
import re
import sys


def main():
    sass_csv, dis, kname, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
    rows = list(csv.reader(open(sass_csv)))
    sect, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": [], "hdr": None}
            sect.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and r:
            cur["rows"].append(r)
    s = [x for x in sect if kname in x["name"]][0]
    h = s["hdr"]
    ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    inst = [(r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0), [int(r[i] or 0) for i, _ in stall_cols]) for r in s["rows"]]
    # disassembly: instruction order + line annotations
    lines, cur_line, in_fn = [], None, False
    for ln in open(dis):
        if ln.startswith(".text."):
            in_fn = mangled in ln
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]+\*/", ln):
            lines.append(cur_line)
    n = min(len(lines), len(inst))
    if len(lines) != len(inst):
        sys.stderr.write("warning: %d instructions in the report, %d in the disassembly\n" % (len(inst), len(lines)))
    agg = {}
    for k in range(n):
        key = lines[k]
        a = agg.setdefault(key, [0, 0, [0] * len(stall_cols)])
        a[0] += inst[k][1]
        a[1] += inst[k][2]
        for j, v in enumerate(inst[k][3]):
            a[2][j] += v
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print("kernel: %s\ntotal samples %d, warp instructions %d" % (s["name"], tot, toti))
    allst = [0] * len(stall_cols)
    for a in agg.values():
        for j, v in enumerate(a[2]):
            allst[j] += v
    print("stall totals:", ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / tot) for (i, c), v in sorted(zip(stall_cols, allst), key=lambda t: -t[1])[:8]))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(zip([c[6:] for _, c in stall_cols], a[2]), key=lambda t: -t[1])[:3]
        print("%5.1f%% samples %5.1f%% inst  %s:%s  [%s]" % (100.0 * a[0] / tot, 100.0 * a[1] / toti, key[0] if key else "?", key[1] if key else "?",
                                                           ", ".join("%s %d" % t for t in st)))


if __name__ == "__main__":
    main()
