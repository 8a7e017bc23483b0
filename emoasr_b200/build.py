"""Builds emoasr_b200/lib/libemoasr_b200.so in-tree with nvcc for sm_100a (no torch headers: the
library is a plain C-ABI shared object, include/emoasr_b200.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libemoasr_b200.so")
SOURCES = ["api.cu", "rnnt_lattice.cu", "ctc.cu", "joint_f32.cu", "joint_bf16.cu", "joint_bwd_ring.cu", "joint_reduce.cu", "ctc_head.cu", "joint_decode.cu", "proj.cu", "ctc_align.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.sep not in p or os.path.exists(p)):
            return p
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "emoasr_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, out=None, extra=()):
    """out / extra: a differently-flagged copy (tuning builds, e.g. -DEMO_ZC_PROF) next to the product library;
    load it with EMOASR_B200_LIB=<path> (tools/ only)."""
    if out is None and not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objdir = os.path.join(HERE, "lib", "obj" if out is None else "obj_" + os.path.splitext(os.path.basename(out))[0])
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("EMO_NVCC_EXTRA", "").split() + list(extra) + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src} ---\n{log}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    target = LIB if out is None else out
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target] + objs)
    return target


if __name__ == "__main__":
    if "--prof" in sys.argv:   # instrumented copy: clock64 wait accounting + tuning switches
        print(build_library(force=True, verbose="-v" in sys.argv,
                            out=os.path.join(HERE, "lib", "libemoasr_b200_prof.so"), extra=["-DEMO_ZC_PROF"]))
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
