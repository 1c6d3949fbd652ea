"""Build libb2me.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

Usage: python build.py [--force]
The .so lands in ./lib/ (git-ignored, but it travels to the GPU box with the repo snapshot).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb2me.so")
SOURCES = ["api.cu", "coords.cu", "spconv_simt.cu", "spconv_tc.cu", "heads.cu", "cluster.cu", "pose.cu", "pointnet.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-DB2ME_BUILD",
]


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "b2me.h")]
    for f in files:
        with open(os.path.join(CSRC, f), "rb") as fp:
            h.update(f.encode())
            h.update(fp.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    global LIBDIR, LIB
    extra = os.environ.get("B2ME_EXTRA_NVCC_FLAGS", "").split()
    if extra:  # debug variants (e.g. -DB2ME_TC_PROFILE) are built beside the product library, never over it
        LIBDIR = os.path.join(HERE, "lib_debug")
        LIB = os.path.join(LIBDIR, "libb2me.so")
        NVCC_FLAGS.extend(extra)
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libb2me.digest")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            failed = True
    with open(os.path.join(LIBDIR, "build.log"), "w") as fp:
        fp.write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed building libb2me")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as fp:
        fp.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
