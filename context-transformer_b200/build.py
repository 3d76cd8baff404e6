"""Build ``libctx_b200.so`` — the C-ABI CUDA library of the hot path — in-tree for sm_100a.

``python -m context_transformer_b200.build`` or ``build_library()``.  nvcc cross-compiles without a
GPU.  Objects go to ``csrc/_obj/`` (git-ignored), the library next to this file so that it travels
to the GPU box with the source snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libctx_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-diag-suppress', '550']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _headers():
    hs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h')))
    hs.append(os.path.join(HERE, '..', 'include', 'ctx_b200.h'))
    return hs


def _digest(paths, extra=''):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """An object is rebuilt when the CONTENT of its source, of any header or the compiler command changed (a digest stored next
    to the object), not by modification time: a checkout or a copy to another machine neither forces nor hides a rebuild."""
    obj_dir = os.path.join(CSRC, '_obj')
    os.makedirs(obj_dir, exist_ok=True)
    hdr = _digest(_headers(), ' '.join([NVCC] + ARCH + FLAGS))
    jobs, objs, stamps = [], [], []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src[:-3] + '.o')
        objs.append(o)
        want = _digest([s], hdr)
        stamp = o + '.sha256'
        have = open(stamp).read().strip() if os.path.exists(stamp) and os.path.exists(o) else ''
        if force or verbose or have != want:
            jobs.append([NVCC] + ARCH + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o])
            stamps.append((stamp, want))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for r in ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
                if verbose or r.returncode:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode:
                    raise RuntimeError('nvcc failed: ' + ' '.join(r.args))
        for stamp, want in stamps:
            with open(stamp, 'w') as f:
                f.write(want)
    if jobs or not os.path.exists(LIB):
        subprocess.check_call([NVCC] + ARCH + ['-shared', '-o', LIB] + objs)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
