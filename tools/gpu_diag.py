"""Run every GPU kernel case in its own process and record pass/fail + stderr tails (first-contact diagnostics)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_kernels_gpu import CASES, run_case  # noqa: E402

out = {}
only = sys.argv[1:] or CASES
for c in only:
    t0 = time.time()
    try:
        rc, so, se = run_case(c, timeout=240)
    except subprocess.TimeoutExpired:
        rc, so, se = -9, "", "TIMEOUT"
    out[c] = {"rc": rc, "sec": round(time.time() - t0, 1), "stdout": so[-600:], "stderr": se[-1500:]}
    print(c, "OK" if rc == 0 else "FAIL", out[c]["sec"], flush=True)
    if rc != 0:
        print(se[-1500:], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_cases.json"), "w"), indent=1)
