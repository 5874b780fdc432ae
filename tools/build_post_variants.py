"""Variants of post.cu with extra -D switches: tools/build_post_variants.py name=-DX=1,-DY=2 ...  -> seld_b200/build/variants/lib_<name>.so"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seld_b200 import build as B
out_dir = os.path.join(B.HERE, 'build', 'variants')
os.makedirs(out_dir, exist_ok=True)
nvcc = B.find_nvcc()
B.build()
for spec in sys.argv[1:]:
    name, _, defs = spec.partition('=')
    obj = os.path.join(out_dir, f'post_{name}.o')
    r = subprocess.run([nvcc, *B.NVCC_FLAGS, *[d for d in defs.split(',') if d], '-Xptxas', '-v', '-c', os.path.join(B.CSRC, 'post.cu'), '-o', obj],
                       capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr)
    print(name, [l for l in r.stderr.splitlines() if 'stats_partial_kernel' in l or 'registers' in l][:4])
    others = [os.path.join(B.HERE, 'build', s.replace('.cu', '.o')) for s in B.SOURCES if s != 'post.cu']
    lib = os.path.join(out_dir, f'lib_{name}.so')
    subprocess.run([nvcc, '-shared', '-o', lib, obj, *others, '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart'], check=True)
    print('built', lib)
