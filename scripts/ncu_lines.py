"""Per-source-line summary of an ncu report: python scripts/ncu_lines.py <report.ncu-rep> [min_pct]
(reads `ncu --page source --print-source cuda,sass --csv`; needs -lineinfo and --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = ""
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and r[0] not in ("", "Function Name") and len(r) > ci:
        try:
            lines.append((fname, int(r[0]), r[1].strip(), float(r[ci] or 0), float(r[si] or 0), float(r[ti] or 0)))
        except ValueError:
            pass
ti_ = sum(l[3] for l in lines)
ts_ = sum(l[4] for l in lines)
print("total warp inst %.4g samples %.4g" % (ti_, ts_))
for f, ln, src, i, s, t in lines:
    if 100 * i / ti_ >= thr or 100 * s / ts_ >= thr:
        print("%-14s %4d inst %5.1f%% samp %5.1f%% thr/inst %4.1f  %s" % (f, ln, 100 * i / ti_, 100 * s / ts_, t / i if i else 0, src[:100]))
if len(sys.argv) > 3:  # line ranges "a-b,c-d,..." of the main .cu file
    main = max(set(l[0] for l in lines), key=lambda f: sum(l[3] for l in lines if l[0] == f))
    for rg in sys.argv[3].split(","):
        a, b = (int(v) for v in rg.split("-"))
        i = sum(l[3] for l in lines if l[0] == main and a <= l[1] <= b)
        s = sum(l[4] for l in lines if l[0] == main and a <= l[1] <= b)
        t = sum(l[5] for l in lines if l[0] == main and a <= l[1] <= b)
        print("%s %4d-%4d inst %5.1f%% samp %5.1f%% thr/inst %4.1f" % (main, a, b, 100 * i / ti_, 100 * s / ts_, t / i if i else 0))
    for f in sorted(set(l[0] for l in lines)):
        if f != main:
            i = sum(l[3] for l in lines if l[0] == f); s = sum(l[4] for l in lines if l[0] == f)
            print("%s inst %5.1f%% samp %5.1f%%" % (f, 100 * i / ti_, 100 * s / ts_))
