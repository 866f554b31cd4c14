"""Phase timeline of the graphed XE / RL training step: external timing events recorded inside the captured graph
(tape.mark), read after a replay.  python r2_phase_timeline.py xe|rl [dedup]"""
import runpy
import sys

import torch

sys.path.insert(0, '.')
from recurrent_fusion_network_b200 import tape  # noqa: E402

which = sys.argv[1]
dd = sys.argv[2] if len(sys.argv) > 2 else "1"
tape.MARKS = []
sys.argv = ["x", dd, "3"]
runpy.run_path(f"profiles/experiments/r2_{which}_graph_once.py", run_name="__main__")
torch.cuda.synchronize()
marks = list(tape.MARKS)
t0 = marks[0][1]
prev = 0.0
for name, ev in marks:
    t = t0.elapsed_time(ev)
    print(f"{t:9.3f} ms  (+{t - prev:7.3f})  {name}")
    prev = t
