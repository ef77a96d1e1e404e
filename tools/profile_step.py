#!/usr/bin/env python3
"""One warm pass + one profiled pass of the whole hot path (device init + run) for ncu.
usage: profile_step.py [B]; prints WARM_LAUNCHES=<kernel launches of the warm pass>."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import redtime_b200 as rt
from redtime_b200 import workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = rt.RedTimeB200()
h.add_cosmologies(wl.make_cosmologies(B, wl.load_example1(4)))
h.prepare()
h.run_resident()
print("WARM_LAUNCHES=%d" % h.launch_count(), flush=True)
h.device_init()
h.run_resident()
print("TOTAL_LAUNCHES=%d" % h.launch_count(), flush=True)
