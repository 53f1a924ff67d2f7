"""Build a variant of libdagma_b200.so with extra nvcc flags (A/B timing, debug stamps):
    python scripts/build_variant.py trace -DDAGMA_SWEEP_TRACE      ->  build/variants/libdagma_trace.so
Select it at run time with DAGMA_B200_LIB=build/variants/libdagma_trace.so."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import NVCC_FLAGS, CSRC
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", "variants")
obj = os.path.join(out, name)
os.makedirs(obj, exist_ok=True)
procs, objs = [], []
for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
    o = os.path.join(obj, os.path.basename(s)[:-3] + ".o")
    objs.append(o)
    procs.append(subprocess.Popen(["/usr/local/cuda/bin/nvcc", *NVCC_FLAGS, *flags, "-c", s, "-o", o], cwd=ROOT))
assert all(p.wait() == 0 for p in procs)
lib = os.path.join(out, f"libdagma_{name}.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs])
print(lib)
