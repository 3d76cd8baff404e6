"""TEST INFRASTRUCTURE ONLY — build the checker binaries.

  build_c()    gcc oracle/c/nms_oracle.c -> oracle/_build/libnms_oracle.so      (always)
  build_ref()  Cython build of the reference's OWN utils/nms/cpu_nms.pyx, read where it lies under
               /root/reference, with the 4-token numpy-2/Cython-3 patch of SURVEY.md §8c applied to
               a temporary copy under /tmp (np.int_t->np.intp_t x2, np.int->np.intp,
               "np.float thresh"->"float thresh").  Output only into oracle/_ref/ (git-ignored,
               travels to the GPU box).  No reference source is stored in the repo.
Both are invoked by __graft_entry__.build(); build_ref() is skipped when /root/reference is absent.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('CTX_REFERENCE_ROOT', '/root/reference')


def build_c(force=False):
    out_dir = os.path.join(HERE, '_build')
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(HERE, 'c', 'nms_oracle.c')
    out = os.path.join(out_dir, 'libnms_oracle.so')
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-o', out, src, '-lm'])
    return out


def build_ref(force=False):
    pyx = os.path.join(REF_ROOT, 'utils', 'nms', 'cpu_nms.pyx')
    if not os.path.isfile(pyx):
        return None
    out_dir = os.path.join(HERE, '_ref')
    os.makedirs(out_dir, exist_ok=True)
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    out = os.path.join(out_dir, 'cpu_nms' + ext)
    if not force and os.path.exists(out):
        return out
    import numpy as np
    tmp = tempfile.mkdtemp(prefix='ctx_ref_nms_')
    try:
        text = open(pyx).read()
        text = text.replace('np.int_t', 'np.intp_t').replace('dtype=np.int)', 'dtype=np.intp)')
        text = text.replace('np.float thresh', 'float thresh')
        text = text.replace('\t', '        ')      # the file mixes tabs into comment lines
        with open(os.path.join(tmp, 'cpu_nms.pyx'), 'w') as f:
            f.write(text)
        subprocess.check_call([sys.executable, '-m', 'cython', '-3', 'cpu_nms.pyx'], cwd=tmp)
        inc = sysconfig.get_paths()['include']
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-w', '-I', inc, '-I', np.get_include(),
                               '-o', out, 'cpu_nms.c'], cwd=tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


if __name__ == '__main__':
    print(build_c(force=True))
    print(build_ref(force=True))
