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


def _fingerprint(src):
    """sha256 over the source file, every header it can include and the compiler command: an object is reused only if this
    matches the stamp written next to it (modification times do not survive checkouts / container snapshots reliably)."""
    import hashlib
    h = hashlib.sha256()
    h.update(' '.join([NVCC] + ARCH_FLAGS + CFLAGS).encode())
    paths = [os.path.join(CSRC, src)]
    for root in (CSRC, os.path.join(REPO, 'include')):
        paths += sorted(os.path.join(root, f) for f in os.listdir(root) if f.endswith(('.cuh', '.h')))
    for path in paths:
        h.update(os.path.basename(path).encode())
        with open(path, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _compile_one(src, verbose):
    obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
    spath = os.path.join(CSRC, src)
    stamp, fp = obj + '.sha256', _fingerprint(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return obj, ''
    cmd = [NVCC] + ARCH_FLAGS + CFLAGS + ['-c', spath, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}\n{}'.format(src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    with open(obj + '.ptxas.log', 'w') as f:
        f.write(r.stderr)
    with open(stamp, 'w') as f:
        f.write(fp)
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
    # the library is re-linked whenever the set of object fingerprints differs from the one it was linked from
    link_fp = '\n'.join(open(o + '.sha256').read().strip() for o in objs)
    link_stamp = LIB_PATH + '.sha256'
    if force or not os.path.exists(LIB_PATH) or not os.path.exists(link_stamp) or open(link_stamp).read() != link_fp:
        cmd = [NVCC] + ARCH_FLAGS + ['-shared', '-o', LIB_PATH] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n{}\n{}'.format(r.stdout, r.stderr))
        with open(link_stamp, 'w') as f:
            f.write(link_fp)
    return LIB_PATH


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose=True))
