"""
Builds libamtfeat.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo
snapshot to the GPU box).  `python amt_tools_b200/build.py` or __graft_entry__.build().
"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libamtfeat.so')
SOURCES = ['api.cpp', 'host_plan.cpp', 'kernels.cu', 'ingest.cu']
HEADERS = ['plan.h', 'fft_device.cuh', 'device_guard.h', os.path.join('..', '..', 'include', 'amtfeat.h')]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, out=None, defines=()):
    """`out` / `defines` build an A/B variant (tools/variants.py) next to the product library."""
    if out is None and not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    target = out or LIB
    objdir = os.path.join(HERE, 'build', os.path.basename(target) + '.obj')
    os.makedirs(objdir, exist_ok=True)
    common = [_nvcc(), '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC,-fvisibility=hidden',
              '-gencode', 'arch=compute_100a,code=sm_100a', '-x', 'cu', '-Xcompiler', '-DAMTFEAT_BUILD'] + ['-D' + d for d in defines]
    if verbose:
        common += ['-Xptxas', '-v']

    def compile_one(src):
        obj = os.path.join(objdir, src + '.o')
        res = subprocess.run(common + ['-c', os.path.join(CSRC, src), '-o', obj], capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError('nvcc failed compiling %s' % src)
        if verbose:
            sys.stderr.write(res.stderr)
    res = subprocess.run([_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', target] + [o for _, o, _ in results],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed linking libamtfeat.so')
    return target


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
