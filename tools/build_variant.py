"""Builds libpg_b200.so variants with extra -D flags into variants/<name>.so (kernel A/B experiments; load one with
PG_B200_LIB=variants/<name>.so).  Usage: python tools/build_variant.py name -DPG_EPI_GROUPS=3 ..."""
import os, subprocess, sys, tempfile, concurrent.futures
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_generative_b200 import _build as B

name, flags = sys.argv[1], sys.argv[2:]
tmp = tempfile.mkdtemp(prefix="pgvar_")
def cc(src):
    obj = os.path.join(tmp, src.replace(".cu", ".o"))
    subprocess.run([B._nvcc(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c", os.path.join(B.CSRC, src), "-o", obj], check=True)
    return obj
with concurrent.futures.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, B.SOURCES))
out = os.path.join(ROOT, "variants", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.run([B._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print("built", out)
