"""Builds libmarie_b200.so (sm_100a) in-tree with nvcc.  Called by __graft_entry__.build().

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmarie_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    with open(path, "rb") as f:
        h.update(f.read())
    for hdr in sorted(os.listdir(CSRC)):
        if hdr.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, hdr), "rb") as f:
                h.update(f.read())
    with open(os.path.join(HERE, "..", "include", "marie_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    stamp = obj + ".sha1"
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    extra = []
    with open(path) as f:
        first = f.readline()
    if first.startswith("// NVCC_FLAGS:"):
        extra = first.split(":", 1)[1].split()
    cmd = ["nvcc", *NVCC_FLAGS, *extra, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, r.stderr


def build(verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs,
               "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
