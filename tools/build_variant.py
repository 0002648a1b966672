"""Build an experiment variant of libwae_b200.so: wn_stack_bf16.cu (or another source) recompiled with extra -D switches, the other
objects taken from the regular build.  The variant lands in wavenet_autoencoders_b200/variants/libwae_<name>.so (git-ignored, travels
with gpurun) and is loaded by the timing tools through WAE_LIB_VARIANT=<name>; the product always loads libwae_b200.so.

    python tools/build_variant.py <name> [-DFOO=1 ...] [--src path/to/alternative/wn_stack_bf16.cu]
"""
import subprocess, sys, os
from pathlib import Path
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavenet_autoencoders_b200 import build as B

name = sys.argv[1]
defs = [a for a in sys.argv[2:] if a.startswith("-D")]
src = B.CSRC / "wn_stack_bf16.cu"
if "--src" in sys.argv:
    src = Path(sys.argv[sys.argv.index("--src") + 1])
B.build()
out = B.PKG / "variants"; out.mkdir(exist_ok=True)
obj = out / f"wn_stack_bf16_{name}.o"
subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *defs, "-I", str(B.CSRC), "-c", str(src), "-o", str(obj)], check=True)
objs = [str(obj if s == "wn_stack_bf16.cu" else B.OBJ / s.replace(".cu", ".o")) for s in B.SOURCES]
lib = out / f"libwae_{name}.so"
subprocess.run([B._nvcc(), "-shared", "--cudart", "static", "-o", str(lib), *objs, "-lpthread", "-ldl", "-lrt"], check=True)
obj.unlink()
print(lib)
