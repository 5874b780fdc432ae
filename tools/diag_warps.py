"""Diagnostic: extract time vs resident warps per SM (SELD_WARPS), to tell latency-bound from throughput-bound."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for w in (2, 4, 6, 8):
    env = dict(os.environ, SELD_WARPS=str(w))
    r = subprocess.run([sys.executable, os.path.join(HERE, 'time_extract.py'), sys.argv[1] if len(sys.argv) > 1 else 'foa'],
                       env=env, capture_output=True, text=True)
    for line in r.stdout.splitlines():
        if ' 600 ' in line:
            print(f'warps {w}: {line}', flush=True)
