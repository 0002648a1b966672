import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ar_time import run
run(32, 12800, "bf16", "mma", 16, 8)
run(32, 12800, "bf16", "mma", 8, 8)
run(32, 12800, "bf16", "mma", 8, 4)
