"""Figures of one `ncu --set full` capture for bench.py's roofline.traffic / issue:
    python scripts/ncu_to_json.py <report.ncu-rep> <key> <frames per launch> "<capture note>" <csrc file> [...]
updates profiles/r2_ncu.json[key] with the DRAM bytes and warp instructions per frame, the issue-slot utilisation and a hash
of the kernel's source files (bench.py refuses the figures when the sources have changed since)."""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, key, frames, note = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
files = sys.argv[5:]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, u, v = rows[0], rows[1], rows[2]


def val(name):
    i = h.index(name)
    x = float(v[i].replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u[i], 1.0)
    return x * scale


sha = hashlib.sha256()
for f in files:
    with open(os.path.join(ROOT, "polarcub_b200", "csrc", f), "rb") as fh:
        sha.update(fh.read())
entry = {
    "capture": note, "report": os.path.basename(rep), "kernel": v[h.index("Kernel Name")], "frames_per_launch": frames,
    "gpu_time_ms": val("gpu__time_duration.sum") / (1e6 if u[h.index("gpu__time_duration.sum")] == "ns" else 1e3 if u[h.index("gpu__time_duration.sum")] == "us" else 1.0),
    "dram_bytes_per_frame": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / frames,
    "warp_inst_per_frame": val("smsp__inst_executed.sum") / frames,
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": val("launch__registers_per_thread"),
    "l2_hit_rate_pct": val("lts__t_sector_hit_rate.pct"),
    "source_files": files, "source_sha": sha.hexdigest()[:16],
}
path = os.path.join(ROOT, "profiles", "r2_ncu.json")
try:
    with open(path) as f:
        d = json.load(f)
except Exception:
    d = {}
d[key] = entry
with open(path, "w") as f:
    json.dump(d, f, indent=1, sort_keys=True)
print(json.dumps(entry, indent=1))
