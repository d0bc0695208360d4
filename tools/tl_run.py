"""Development aid: run one IAF step with the -DIAF_TC_TIMELINE build of the library and print
the in-kernel timeline of CTA 0 (see TL() in iaf_b200/csrc/iaf_tc.cu)."""
import ctypes
import os
import sys

sys.path.insert(0, os.getcwd())
import iaf_b200.build as B

B.LIB = os.path.join(os.getcwd(), "iaf_b200", "lib", "libiaf_tl.so")
import iaf_b200._lib as L

L.LIB = B.LIB
import torch
from bench import make_workload

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "c2a"
op, layers, sets = make_workload(name, dev, 2)
s = sets[0]
for i in range(3):
    op.step(s["z"], s["ctx"])
torch.cuda.synchronize()
lib = ctypes.CDLL(B.LIB)
lib.iaf_tc_timeline_dump()  # discard warm-up events
print("=== timed launch ===", flush=True)
op.step(s["z"], s["ctx"])
torch.cuda.synchronize()
lib.iaf_tc_timeline_dump()
