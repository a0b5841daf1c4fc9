"""Builds libtt_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python transform-and-tell_b200/csrc/build.py [--force] [--verbose]

Objects are cached per source (rebuilt when the source or a header is newer).  nvcc
cross-compiles without a GPU.  The .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, 'build')
LIB = os.path.join(os.path.dirname(HERE), 'tell_b200', 'libtt_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v',
         '-I', os.path.join(ROOT, 'include')]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith('.cu'))


def headers_mtime():
    m = 0.0
    for d in (HERE, os.path.join(ROOT, 'include')):
        for f in os.listdir(d):
            if f.endswith(('.h', '.cuh')):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def compile_one(src, force, verbose):
    obj = os.path.join(BUILD, src[:-3] + '.o')
    sp = os.path.join(HERE, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(sp), headers_mtime())):
        return obj, ''
    cmd = [NVCC] + FLAGS + ['-c', sp, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    log = r.stderr if verbose else ''
    with open(obj + '.ptxas.log', 'w') as f:
        f.write(r.stderr)
    return obj, log


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: compile_one(s, force, verbose), srcs))
    objs = [o for o, _ in results]
    for _, log in results:
        if log:
            print(log)
    if (force or not os.path.exists(LIB)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    lib = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(lib)
