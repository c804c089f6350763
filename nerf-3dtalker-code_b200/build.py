"""Builds libheadnerf_b200.so (sm_100a) in-tree with nvcc.  No JIT cache, no torch extension: the
product boundary is a plain C-ABI shared library (include/headnerf_b200.h) loaded with ctypes."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libheadnerf_b200.so")
SOURCES = ["hn_api.cu", "hn_composite.cu", "hn_mlp_sched.cu", "hn_mlp_pack.cu", "hn_mlp_fwd.cu",
           "hn_mlp_bwd.cu", "hn_mlp_wgrad.cu", "hn_precise.cu", "hn_fold.cu", "hn_render2d.cu", "hn_render.cu", "hn_train.cu", "hn_fine.cu", "hn_nr.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-DHN_BUILDING_DSO"]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "headnerf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link the shared library next to this file."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = ["-DHN_PIPE_COUNTERS"] if os.environ.get("HN_PIPE_COUNTERS", "0") == "1" else []   # diagnostic build
    objdir = os.path.join(ROOT, "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.isfile(path):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose:
            print(out)
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
