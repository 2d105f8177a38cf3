"""Builds ipavsr_b200/libipavsr_b200.so from csrc/*.cu for sm_100a with nvcc (in-tree; the .so travels to the GPU
box with the snapshot).  `python -m ipavsr_b200.build` or `__graft_entry__.build()`."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libipavsr_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _stale(objs_src):
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = objs_src + glob.glob(os.path.join(SRC, '*.cuh')) + [os.path.join(HERE, '..', 'include', 'ipavsr_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(SRC, '*.cu')))
    if not force and not _stale(srcs):
        return OUT
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(
                [os.path.getmtime(s)] + [os.path.getmtime(h) for h in glob.glob(os.path.join(SRC, '*.cuh'))] +
                [os.path.getmtime(os.path.join(HERE, '..', 'include', 'ipavsr_b200.h'))]):
            continue
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
