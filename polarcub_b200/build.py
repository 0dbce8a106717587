"""Builds polarcub_b200/libpolarcub_b200.so (the C-ABI library of hand-written sm_100a kernels) with nvcc.

In-tree build so the .so travels with the repo snapshot to the GPU box.  `-fmad=false` because the
reference rounds every product and sum separately (float64); the kernels additionally use the explicit
__dmul_rn / __dadd_rn intrinsics on the parity-critical expressions.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libpolarcub_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--shared", "-Xptxas", "-v",
] + os.environ.get("PC_NVCC_EXTRA", "").split()  # experiment builds only (e.g. -DSCLP_SELECT_SHFL=0)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def stale():
    if not os.path.isfile(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "polarcub_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    only = [t for t in os.environ.get("PC_BUILD_ONLY", "").split(",") if t]  # experiment builds: recompile these, reuse the other objects
    for s in sources():
        o = os.path.join(HERE, "build", os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if only and os.path.basename(s)[:-3] not in only and os.path.isfile(o):
            continue
        cmd = [NVCC] + [f for f in FLAGS if f != "--shared"] + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % s)
    link = [NVCC, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
