"""Build a variant of the library with extra -D flags for A/B kernel experiments on the GPU box:
    python tools/build_variant.py NAME -DT4_MINBLOCKS=10 ...   ->  envgs_b200/_variants/lib_NAME.so
    ENVGS_B200_LIB=envgs_b200/_variants/lib_NAME.so python tools/trace_bench.py"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from envgs_b200 import build as B
name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.HERE, "_variants"); obj_dir = os.path.join(out_dir, name)
os.makedirs(obj_dir, exist_ok=True)
def one(src):
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
    if r.returncode: raise SystemExit(r.stderr)
    open(obj + ".log", "w").write(r.stderr)
    return obj
with ThreadPoolExecutor(8) as ex: objs = list(ex.map(one, B.SOURCES))
lib = os.path.join(out_dir, f"lib_{name}.so")
r = subprocess.run([B._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], capture_output=True, text=True)
if r.returncode: raise SystemExit(r.stderr)
print(lib)
