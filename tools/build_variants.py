"""Build variants of libseld_b200.so with extra -D switches (kernel experiments): tools/build_variants.py name=-DX=1,-DY=0 ...
The libraries land in seld_b200/build/variants/lib_<name>.so; tools/time_variants.py times each of them."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seld_b200 import build as B  # noqa: E402

out_dir = os.path.join(B.HERE, 'build', 'variants')
os.makedirs(out_dir, exist_ok=True)
nvcc = B.find_nvcc()
B.build()                                                   # the other objects
procs = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition('=')
    obj = os.path.join(out_dir, f'extract_{name}.o')
    cmd = [nvcc, *B.NVCC_FLAGS, *[d for d in defs.split(',') if d], '-Xptxas', '-v', '-c', os.path.join(B.CSRC, 'extract.cu'), '-o', obj]
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, obj, p in procs:
    out, _ = p.communicate()
    if p.returncode:
        raise SystemExit(f'{name}: nvcc failed\n{out}')
    with open(os.path.join(out_dir, f'ptxas_{name}.log'), 'w') as fh:
        fh.write(out)
    others = [os.path.join(B.HERE, 'build', s.replace('.cu', '.o')) for s in B.SOURCES if s != 'extract.cu']
    lib = os.path.join(out_dir, f'lib_{name}.so')
    subprocess.run([nvcc, '-shared', '-o', lib, obj, *others, '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart'], check=True)
    print('built', lib)
