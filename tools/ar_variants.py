"""A/B timing of libwae_b200.so variants built into wavenet_autoencoders_b200/build/variants/lib_<name>.so (GPU box only):
every variant is copied over the library and timed in its own process."""
import glob, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wavenet_autoencoders_b200", "libwae_b200.so")
CODE = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import ar_time; "
        "ar_time.run(32, 1280, 'bf16', 'mma', 8, 8); ar_time.run(32, 1280, 'bf16', 'mma', 16, 8)") % (ROOT, os.path.join(ROOT, "tools"))
shutil.copy(LIB, LIB + ".orig")
try:
    for v in sorted(glob.glob(os.path.join(ROOT, "wavenet_autoencoders_b200", "build", "variants", "lib_*.so"))):
        shutil.copy(v, LIB)
        r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, timeout=300)
        print(os.path.basename(v), flush=True)
        print("".join(l + "\n" for l in r.stdout.splitlines() if "us/step" in l) or r.stderr[-800:], flush=True)
finally:
    shutil.move(LIB + ".orig", LIB)
