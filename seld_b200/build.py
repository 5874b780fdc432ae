"""Build libseld_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m seld_b200.build [--force]

The library is the C-ABI boundary declared in include/seld_b200.h.  It is built next to this
file so that it travels with the source tree (no JIT cache, no site-packages install).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libseld_b200.so')
SOURCES = ['extract.cu', 'post.cu', 'mask.cu', 'augment.cu', 'spec_ops.cu', 'gcc_gemm.cu']
HEADERS = ['seld_common.cuh', 'extract_core.cuh', 'mel_pieces.h', 'plan.h', 'philox.cuh', os.path.join('..', '..', 'include', 'seld_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def find_nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: seld_b200 has no prebuilt or CPU fallback path')
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    objs = []
    build_dir = os.path.join(HERE, 'build')
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-Xptxas', '-v', '-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f'$ {" ".join(cmd)}\n{out}')
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
        objs.append(obj)
    with open(os.path.join(build_dir, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    tmp = LIB_PATH + '.tmp'
    cmd = [nvcc, '-shared', '-o', tmp, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(f'built {LIB_PATH}')
    return LIB_PATH


if __name__ == '__main__':
    build(force='--force' in sys.argv)
