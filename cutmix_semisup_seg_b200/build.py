"""In-tree build of libb200seg.so (hand-written sm_100a kernels + C ABI, include/b200seg.h).

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
OBJ_DIR = os.path.join(REPO, 'build', 'obj')
LIB_PATH = os.path.join(PKG_DIR, 'libb200seg.so')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
          '-Xptxas', '-v', '-I', os.path.join(REPO, 'include')]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(REPO, 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cuh', '.h')):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile_one(src, verbose):
    obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
    spath = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(spath), _deps_mtime()):
        return obj, ''
    cmd = [NVCC] + ARCH_FLAGS + CFLAGS + ['-c', spath, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}\n{}'.format(src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    with open(obj + '.ptxas.log', 'w') as f:
        f.write(r.stderr)
    return obj, r.stderr


def build_lib(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libb200seg.so next to this file."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    if force:
        for s in srcs:
            o = os.path.join(OBJ_DIR, s[:-3] + '.o')
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = [o for o, _ in ex.map(lambda s: _compile_one(s, verbose), srcs)]
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = [NVCC] + ARCH_FLAGS + ['-shared', '-o', LIB_PATH] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n{}\n{}'.format(r.stdout, r.stderr))
    return LIB_PATH


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose=True))
