"""Builds libkvfe.so (sm_100a) in-tree with nvcc.  No torch, no JIT cache: the .so sits next to the
sources so that it travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libkvfe.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["api.cu", "pipeline.cu", "rectify.cu", "pyramid.cu", "lk.cu", "gftt.cu", "select.cu", "stereo.cu", "ransac.cu", "fsm.cu", "mesh.cu", "ingest.cu", "rgbd.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-fmad=false",                       # IEEE op-by-op arithmetic; FMAs only where written
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-Xptxas", "-v"]


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            log = os.path.join(OBJ, src + ".log")
            with open(log, "w") as fh:
                fh.write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % src)
            if verbose:
                sys.stderr.write(r.stderr)
            objs.append(obj)
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
