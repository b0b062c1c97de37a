"""Build libradiobear_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m radiobear_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libradiobear_b200.so')
SOURCES = ['alpha_kernels.cu', 'rt_kernels.cu', 'probe_kernels.cu', 'capi.cu']
HEADERS = [os.path.join(CSRC, 'rb_common.cuh'), os.path.join(os.path.dirname(HERE), 'include', 'radiobear_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return 'nvcc'


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=True):
    """Compile every CUDA source into one shared library; returns its path."""
    if not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
