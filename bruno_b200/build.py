"""Builds bruno_b200/lib/libbruno_sm100.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Run as `python -m bruno_b200.build` (or through `__graft_entry__.build()`).  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libbruno_sm100.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

SOURCES = ['runtime.cu', 'process.cu', 'flow_misc.cu', 'dense.cu', 'gemm_tf32.cu', 'conv_sm100.cu', 'coupling_conv.cu', 'optim.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']
NVCC_FLAGS += os.environ.get('BRUNO_NVCC_EXTRA', '').split()   # experiments: e.g. -DBRUNO_EXP_...


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, 'bruno_sm100.h')]
    for f in files:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if os.path.isfile(p):
            with open(p, 'rb') as fh:
                h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'libbruno_sm100.stamp')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src).replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + ['-I', INCLUDE, '-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append('== %s ==\n%s' % (os.path.basename(src), out))
        if p.returncode != 0:
            sys.stderr.write('\n'.join(log))
            raise RuntimeError('nvcc failed on %s' % src)
    with open(os.path.join(LIBDIR, 'ptxas.log'), 'w') as fh:
        fh.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    with open(stamp, 'w') as fh:
        fh.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
