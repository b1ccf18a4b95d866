"""One workload, few launches: the command ncu wraps (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import perf_probe_lib as L

case = sys.argv[1] if len(sys.argv) > 1 else "c3plain"
L.run_case(case, steps=2, warmup=1)
