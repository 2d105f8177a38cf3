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


def source_hash():
    """Hash of everything the library is built from; the .so exports the value it was built with (ipavsr_source_hash)."""
    import hashlib
    h = hashlib.sha1()
    for f in sorted(glob.glob(os.path.join(SRC, '*.cu')) + glob.glob(os.path.join(SRC, '*.cuh'))) + [
            os.path.join(HERE, '..', 'include', 'ipavsr_b200.h')]:
        h.update(os.path.basename(f).encode())
        h.update(open(f, 'rb').read())
    return h.hexdigest()[:16]


def built_hash():
    """The source hash recorded next to the .so by the last build (None when there is none)."""
    try:
        return open(OUT + '.hash').read().strip() if os.path.exists(OUT) else None
    except OSError:
        return None


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(SRC, '*.cu')))
    want = source_hash()
    if not force and built_hash() == want:
        return OUT
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        is_rt = os.path.basename(s) == 'runtime.cu'          # carries the source hash: rebuilt with every build
        if not force and not is_rt and os.path.exists(o) and os.path.getmtime(o) > max(
                [os.path.getmtime(s)] + [os.path.getmtime(h) for h in glob.glob(os.path.join(SRC, '*.cuh'))] +
                [os.path.getmtime(os.path.join(HERE, '..', 'include', 'ipavsr_b200.h'))]):
            continue
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + (
            ['-DIPAVSR_SRC_HASH="%s"' % want] if is_rt else []) + ['-c', s, '-o', o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(cmd)
    with open(OUT + '.hash', 'w') as f:
        f.write(want + '\n')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
