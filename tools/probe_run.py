"""Development aid: run IAF steps with the -DIAF_FZ_PROBE build (libiaf_probe.so) and print where CTA 1's lead lanes
spent their cycles (see PROBE() in iaf_b200/csrc/iaf_fz.cuh).  The numbers are of the LAST launch."""
import ctypes
import os
import sys

sys.path.insert(0, os.getcwd())
import iaf_b200.build as B

B.LIB = os.path.join(os.getcwd(), "iaf_b200", "lib", "libiaf_probe.so")
import iaf_b200._lib as L

L.LIB = B.LIB
import torch
from bench import make_workload

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "c2a"
op, layers, sets = make_workload(name, dev, 4)
for i in range(20):
    s = sets[i % 4]
    op.step(s["z"], s["ctx"])
torch.cuda.synchronize()
ctypes.CDLL(B.LIB).iaf_fz_probe_dump()
