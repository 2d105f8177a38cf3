"""Step-by-step bring-up of the tcgen05 GEMM: each case in a subprocess with a timeout so a hung kernel cannot
take the whole run down."""
import subprocess, sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
CASE = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))
import numpy as np
from test_gpu_gemm_tc import _run
mode, ta, tb, M, N, K = [int(x) for x in sys.argv[1:7]]
print('mode %%d ta %%d tb %%d M %%d N %%d K %%d err %%.3e' %% (mode, ta, tb, M, N, K, _run(mode, ta, tb, M, N, K)), flush=True)
''' % (os.path.dirname(HERE), os.path.dirname(HERE))
cases = []
for mode in (1,):
    for ta, tb in ((0, 0), (1, 0)):
        for (M, N, K) in ((2048, 52, 500), (1040, 128, 1200), (1040, 2000, 1200), (1200, 2000, 4096), (1200, 2000, 20480)):
            cases.append((mode, ta, tb, M, N, K))
for c in cases:
    try:
        r = subprocess.run([sys.executable, '-c', CASE] + [str(x) for x in c], capture_output=True, text=True, timeout=90)
        out = (r.stdout.strip().splitlines() or ['(no output)'])[-1]
        if r.returncode != 0:
            out += ' | rc=%d %s' % (r.returncode, (r.stderr.strip().splitlines() or [''])[-1][:200])
    except subprocess.TimeoutExpired:
        out = 'TIMEOUT %r' % (c,)
    print(out, flush=True)
