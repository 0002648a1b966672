"""AR synthesis at 32 utterances per GPU: clusters x utterances-per-cluster sweep (how many SMs the work is spread over). GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ar_time import run
for cluster, U in [(16, 8), (16, 4), (16, 2), (8, 8), (8, 4), (8, 2)]:
    run(32, 2560, "bf16", "mma", cluster, U)
run(64, 2560, "bf16", "mma", 16, 8)
run(64, 2560, "bf16", "mma", 8, 4)
