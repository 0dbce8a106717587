"""python scripts/launch_summary.py gpurun_out/launches_x.csv -> markdown table of kernels by total time"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        hdr = r
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    v = float(r[vi].replace(",", ""))
    tot[r[ki][:110]] += v; cnt[r[ki][:110]] += 1
s = sum(tot.values())
print("| kernel | launches | total ns | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:14]:
    print("| %s | %d | %.0f | %.1f%% |" % (k, cnt[k], v, 100 * v / s))
