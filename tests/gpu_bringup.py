#!/usr/bin/env python
"""Run every GPU test node in its own process with a timeout, so that a device-side trap or a hang in one kernel
does not poison the CUDA context of the others. Usage: python tests/gpu_bringup.py [pytest -k expr] [file ...]"""
import subprocess
import sys
import time

files = [a for a in sys.argv[1:] if a.endswith(".py")] or ["tests/test_kernels_gpu.py"]
kexpr = [a for a in sys.argv[1:] if not a.endswith(".py")]
cmd = [sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + files + (["-k", kexpr[0]] if kexpr else [])
nodes = [l.strip() for l in subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines() if "::" in l]
print(f"{len(nodes)} test nodes")
fails = 0
for n in nodes:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "--no-header", "-m", "gpu", n],
                           capture_output=True, text=True, timeout=180)
        ok = r.returncode == 0
        tail = "" if ok else "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    except subprocess.TimeoutExpired:
        ok, tail = False, "TIMEOUT"
    fails += 0 if ok else 1
    print(f"[{'PASS' if ok else 'FAIL'}] {n} ({time.time() - t0:.1f}s)")
    if tail:
        print("    " + tail.replace("\n", "\n    "))
    sys.stdout.flush()
print(f"{fails} failed of {len(nodes)}")
sys.exit(1 if fails else 0)
