"""In-tree build of libfpc_b200.so (sm_100a only).  `python -m fpc_diffrend_b200.build [--force] [-v]`"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libfpc_b200.so')
NVCC = os.environ.get('FPC_NVCC', '/usr/local/cuda/bin/nvcc')
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hdrs.append(os.path.join(HERE, '..', 'include', 'fpc_b200.h'))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose, defines=(), tag=''):
    obj = src[:-3] + tag + '.o'
    if not _stale(obj, [src] + _deps()):
        return obj
    cmd = [NVCC] + ARCH_FLAGS + CFLAGS + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed on %s' % src)
    return obj


def build(force=False, verbose=False, defines=(), tag=''):
    """Compile csrc/*.cu and link the library.  `defines` / `tag` build a kernel-variant copy
    libfpc_b200<tag>.so (experiments only; the product is the untagged library)."""
    srcs = sources()
    lib = LIB if not tag else LIB[:-3] + tag + '.so'
    if force:
        for s in srcs:
            o = s[:-3] + tag + '.o'
            if os.path.exists(o):
                os.remove(o)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, defines, tag), srcs))
    if force or _stale(lib, objs):
        cmd = [NVCC] + ARCH_FLAGS + ['-shared', '-o', lib] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
