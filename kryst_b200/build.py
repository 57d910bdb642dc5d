"""In-tree build of libkryst_b200.so (hand-written sm_100a CUDA kernels + C ABI).

nvcc cross-compiles without a GPU.  Objects go to kryst_b200/build/, the library to
kryst_b200/libkryst_b200.so (git-ignored, but it travels to the GPU box with the snapshot).
-fmad=false is REQUIRED: the kernels promise the oracle's exact mul/add sequence.
"""
import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(os.path.dirname(HERE), "include")
OUT = os.path.join(HERE, "libkryst_b200.so")
OBJ = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-DKB_NO_FMA",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-I", INC, "-I", CSRC]


def _digest(paths):
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def build(verbose=False, force=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(INC, "*.h")))
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp.txt")
    dig = _digest(srcs + hdrs)
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    hd = _digest(hdrs)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        tag = obj + ".tag"
        d = hashlib.sha256((hd + open(src).read()).encode()).hexdigest()
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == d:
            return obj
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        open(tag, "w").write(d)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, "-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++", "-o", OUT] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
