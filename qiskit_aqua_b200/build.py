"""Builds libpsum.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

psum.cu holds the host logic and the small kernels; every (register bits, threads) configuration of
the hot kernel k_pass is its own translation unit (psum_kpass_inst.cu with -DPSUM_INST_RB/-DPSUM_INST_NT)
so that they compile in parallel."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libpsum.so')
OBJDIR = os.path.join(CSRC, 'build')
KPASS_CONFIGS = [(3, 256), (3, 512), (4, 128), (4, 256)]
DEPS = ['psum.cu', 'psum_kpass_inst.cu', 'psum_kpass.h', 'psum_kernels.cuh', 'psum_plan.h', 'psum_eig.h',
        'psum_lanczos.inc', 'psum_hostcopy.h', os.path.join('..', '..', 'include', 'psum.h')]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    return res.returncode, res.stdout + res.stderr


def build_lib(force=False, verbose=False, lib_path=None, extra_flags=()):
    lib_path = lib_path or LIB
    if not force and lib_path == LIB and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(OBJDIR, exist_ok=True)
    common = [nvcc, '-O3', '-std=c++17', *ARCH, '-lineinfo', '-Xcompiler', '-fPIC']
    common += os.environ.get('PSUM_NVCC_FLAGS', '').split() + list(extra_flags)
    if verbose:
        common += ['-Xptxas', '-v']
    tag = os.path.basename(lib_path)
    jobs = [(common + ['-c', os.path.join(CSRC, 'psum.cu'), '-o', os.path.join(OBJDIR, tag + '.psum.o')])]
    for rb, nt in KPASS_CONFIGS:
        jobs.append(common + ['-DPSUM_INST_RB=%d' % rb, '-DPSUM_INST_NT=%d' % nt, '-c',
                              os.path.join(CSRC, 'psum_kpass_inst.cu'), '-o',
                              os.path.join(OBJDIR, '%s.kpass_%d_%d.o' % (tag, rb, nt))])
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        results = list(ex.map(_run, jobs))
    log = ''.join(out for _, out in results)
    if any(rc != 0 for rc, _ in results):
        sys.stderr.write(log)
        raise RuntimeError('nvcc failed building libpsum.so')
    objs = [j[-1] for j in jobs]
    rc, out = _run([nvcc, *ARCH, '-shared', '-o', lib_path] + objs + ['-ldl'])
    if rc != 0:
        sys.stderr.write(log + out)
        raise RuntimeError('nvcc failed linking libpsum.so')
    if verbose:
        sys.stderr.write(log + out)
    return lib_path


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv))
