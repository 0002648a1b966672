import sys, os
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
from ar_time import run
for cluster, U in [(16, 2), (16, 4), (16, 1), (8, 4), (8, 2)]:
    try:
        run(32, 1280, "fp32", "simt", cluster, U)
    except Exception as e:
        print("fp32", cluster, U, "failed:", str(e)[:150])
for cluster, U in [(16, 4), (8, 4), (8, 2)]:
    try:
        run(32, 1280, "bf16", "simt", cluster, U)
    except Exception as e:
        print("bf16 simt", cluster, U, "failed:", str(e)[:150])
