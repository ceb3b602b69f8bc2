"""In-tree build of libpdmp_b200.so: nvcc, sm_100a only, -lineinfo (ncu source pages).

    python -m pdmp_b200.build        or        __graft_entry__.build()

One object per model translation unit, compiled in parallel, linked into one shared
library next to this file (it travels to the GPU box with the repo snapshot).
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpdmp_b200.so")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _deps_hash(extra):
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(FLAGS + extra).encode())
    return h.hexdigest()


def _compile(src, obj, extra, log):
    cmd = [NVCC] + FLAGS + extra + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr[-4000:]}")
    return obj


def build(force=False, extra_flags=(), verbose=False, jobs=None):
    extra = list(extra_flags)
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdr = _deps_hash(extra)
    todo, objs = [], []
    for f in srcs:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        stamp = obj + ".stamp"
        key = hdr + hashlib.sha256(open(src, "rb").read()).hexdigest()
        objs.append(obj)
        if force or not os.path.exists(obj) or not os.path.exists(stamp) or open(stamp).read() != key:
            todo.append((src, obj, stamp, key))
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as ex:
            futs = {ex.submit(_compile, s, o, extra, o + ".log"): (o, st, k) for s, o, st, k in todo}
            for fu in cf.as_completed(futs):
                o, st, k = futs[fu]
                fu.result()
                open(st, "w").write(k)
                if verbose:
                    print("compiled", os.path.basename(o), flush=True)
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
