"""Diagnostic: extract time vs consecutive frames per team (SELD_FPW)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
mode = sys.argv[1] if len(sys.argv) > 1 else 'foa'
for f in (1, 2, 3, 4, 8):
    env = dict(os.environ, SELD_FPW=str(f))
    r = subprocess.run([sys.executable, os.path.join(HERE, 'time_extract.py'), mode], env=env, capture_output=True, text=True)
    for line in r.stdout.splitlines():
        if ' 600 ' in line:
            print(f'fpw {f}: {line}', flush=True)
