"""Builds libpsum.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libpsum.so')
SOURCES = ['psum.cu']
DEPS = ['psum.cu', 'psum_kernels.cuh', 'psum_plan.h', 'psum_eig.h', 'psum_lanczos.inc',
        os.path.join('..', '..', 'include', 'psum.h')]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_lib(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
           '-Xcompiler', '-fPIC', '-shared', '-o', LIB]
    cmd += os.environ.get('PSUM_NVCC_FLAGS', '').split()
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ['-ldl']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libpsum.so')
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv))
