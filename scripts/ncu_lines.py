#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
   python tools_ncu_lines.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
out = []
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr[4:], r[4:]))
        try:
            out.append((fname, int(r[0]), r[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"]),
                        {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}))
        except Exception:
            pass
tot_s = sum(o[3] for o in out) or 1
tot_i = sum(o[4] for o in out) or 1
print("total samples", tot_s, "total warp instr", tot_i)
for o in sorted(out, key=lambda o: -o[3])[:top]:
    st = sorted(o[5].items(), key=lambda kv: -kv[1])[:3]
    print("%s:%d  samp %.1f%%  inst %.1f%%  %s | %s" % (o[0], o[1], 100 * o[3] / tot_s, 100 * o[4] / tot_i, st, o[2]))
