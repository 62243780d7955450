"""Sweep the number of co-scheduled CTA pairs of the pair GEMM (CORRIF_PAIR_CLUSTERS) on one big shape."""
import os, subprocess, sys
for n in (74, 72, 70, 68, 66, 64, 60, 56):
    env = dict(os.environ, CORRIF_PAIR_CLUSTERS=str(n))
    out = subprocess.run([sys.executable, "tools/gemm_bench.py"], env=env, capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith(("linear mm qkv", "dgrad mm qkv", "wgrad mm qkv"))]
    print(n, " | ".join(l.split("split")[1].strip() for l in lines), flush=True)
