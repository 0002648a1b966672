"""Run the GPU test-suite one test FUNCTION per process (a trapped kernel poisons its CUDA context, so isolation
keeps one bug from hiding the rest) and write a summary to gpurun_out/.  Usage (on the GPU box):

    python tools/gpu_runner.py [-k substring] [--timeout 600]
"""
import argparse
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default="")
    ap.add_argument("--timeout", type=int, default=600)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_tests.log"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    col = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "--collect-only", "-q"],
                         cwd=ROOT, capture_output=True, text=True)
    funcs = []
    for line in col.stdout.splitlines():
        m = re.match(r"(tests/[\w/]+\.py::\w+)", line)
        if m and m.group(1) not in funcs and a.k in m.group(1):
            funcs.append(m.group(1))
    results = []
    with open(a.out, "w") as log:
        for f in funcs:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, "-m", "pytest", f, "-m", "gpu", "-q", "-x", "--no-header", "-p", "no:cacheprovider"],
                                   cwd=ROOT, capture_output=True, text=True, timeout=a.timeout)
                status, out = ("PASS" if r.returncode == 0 else f"FAIL({r.returncode})"), r.stdout[-6000:] + r.stderr[-3000:]
            except subprocess.TimeoutExpired as e:
                status, out = "TIMEOUT", (e.stdout or b"").decode(errors="replace")[-3000:] if isinstance(e.stdout, bytes) else str(e.stdout)[-3000:]
            dt = time.time() - t0
            results.append((f, status, dt))
            log.write(f"===== {f}: {status} ({dt:.1f}s)\n")
            if not status.startswith("PASS"):
                log.write(out + "\n")
            log.flush()
            print(f"{status:10s} {dt:6.1f}s {f}", flush=True)
    bad = [r for r in results if not r[1].startswith("PASS")]
    print(f"{len(results) - len(bad)}/{len(results)} test functions passed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
