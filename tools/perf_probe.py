"""Quick device-timed throughput probe (development tool, not the bench contract)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import perf_probe_lib as L
for case in (sys.argv[1:] or ["c3", "c3plain", "c3cs", "c4", "c1", "c2", "c2one"]):
    L.run_case(case)
