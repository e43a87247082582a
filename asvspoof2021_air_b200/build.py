"""In-tree build of libair_b200.so (hand-written sm_100a kernels + the C-ABI).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repository snapshot.  `python -m asvspoof2021_air_b200.build` or `__graft_entry__.build()`.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libair_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false",
              "-Xptxas", "-v", "-I", os.path.join(os.path.dirname(HERE), "include"), "-I", CSRC]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libair_b200.so")
    return exe


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src,) + tuple(extra))


def _compile(args):
    src, obj, headers, verbose = args
    if not _newer(src, obj, headers):
        return src, "", False
    if src.endswith(".cpp"):                 # host-only sources (audio ingest): plain C++ through nvcc's host compiler
        cmd = [_nvcc(), "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall", "-I", CSRC,
               "-I", os.path.join(os.path.dirname(HERE), "include"), "-c", src, "-o", obj]
    else:
        flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
        cmd = [_nvcc()] + ARCH + flags + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return src, r.stderr, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))
    headers = tuple(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers += tuple(os.path.join(os.path.dirname(HERE), "include", f)
                     for f in os.listdir(os.path.join(os.path.dirname(HERE), "include")) if f.endswith(".h"))
    objs = [os.path.join(OBJ, os.path.splitext(os.path.basename(s))[0] + ".o") for s in srcs]
    if force:
        for o in objs:
            if os.path.exists(o):
                os.remove(o)
    changed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for src, log, did in ex.map(_compile, [(s, o, headers, verbose) for s, o in zip(srcs, objs)]):
            changed |= did
            if did and verbose:
                print("[build] %s\n%s" % (os.path.basename(src), log))
    if changed or not os.path.exists(LIB):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcuda", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
