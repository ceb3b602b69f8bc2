"""Builds a variant of libpdmp_b200.so with extra nvcc flags into piecewisedeterministicmarkovprocesses.jl_b200/variants/
(kernel A/B experiments; run with PDMP_B200_LIB=<path>).  Usage: [VARIANT_ONLY=inst_tcp] python scripts/build_variant.py NAME -DFOO=1 ..."""
import os, shutil, subprocess, sys, concurrent.futures as cf
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdmp_b200 import build as B
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(B.HERE, "variants"); obj = os.path.join(out, "obj_" + name)
os.makedirs(obj, exist_ok=True)
only = os.environ.get("VARIANT_ONLY", "").split(",") if os.environ.get("VARIANT_ONLY") else None
srcs = sorted(f for f in os.listdir(B.CSRC) if f.endswith(".cu"))
def comp(f):
    o = os.path.join(obj, f[:-3] + ".o")
    base = os.path.join(B.OBJ, f[:-3] + ".o")
    if only and f[:-3] not in only and os.path.exists(base):
        shutil.copy(base, o); return o          # unchanged translation units come from the main build
    r = subprocess.run([B.NVCC] + B.FLAGS + flags + ["-c", os.path.join(B.CSRC, f), "-o", o], capture_output=True, text=True)
    open(o + ".log", "w").write(r.stdout + r.stderr)
    if r.returncode: raise RuntimeError(r.stderr[-3000:])
    return o
with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(comp, srcs))
lib = os.path.join(out, f"libpdmp_b200_{name}.so")
subprocess.check_call([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
print(lib)
