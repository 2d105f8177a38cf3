import sys, os, ctypes as C, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ipavsr_b200 import _lib
from ipavsr_b200.engine import _PackPlan
N, T, F = 512, 40, 1200
rng = np.random.default_rng(0)
lens = rng.integers(12, T + 1, size=N)
plan = _PackPlan(lens, T)
pin = torch.empty(1 << 20, dtype=torch.int32).pin_memory()
plan.upload(torch.device('cuda'), pin)
host = torch.empty(N, T, F, dtype=torch.float32).pin_memory(); host.normal_()
dev = torch.empty(N * T, F, device='cuda')
cs = torch.cuda.Stream()
a = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16); b = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)
# our own GEMM as the compute load
M, K, Nn = 13325, 1200, 2000
A = torch.randn(M, K, device='cuda'); B = torch.randn(K, Nn, device='cuda'); Cm = torch.empty(M, Nn, device='cuda')
ah, al = torch.empty(M, K, dtype=torch.float16, device='cuda'), torch.empty(M, K, dtype=torch.float16, device='cuda')
bh, bl = torch.empty(K, Nn, dtype=torch.float16, device='cuda'), torch.empty(K, Nn, dtype=torch.float16, device='cuda')
sc = torch.zeros(4, device='cuda')
st = lambda s: C.c_void_p(s.cuda_stream)
main = torch.cuda.current_stream()
_lib.call('ipavsr_f16_split', A.data_ptr(), K, M, K, ah.data_ptr(), al.data_ptr(), K, sc.data_ptr(), sc.data_ptr() + 4, 0, st(main))
_lib.call('ipavsr_f16_split', B.data_ptr(), Nn, K, Nn, bh.data_ptr(), bl.data_ptr(), Nn, sc.data_ptr() + 8, sc.data_ptr() + 12, 0, st(main))
def ours(n=40):
    for _ in range(n):
        _lib.call('ipavsr_gemm_f16x3', 0, 0, M, Nn, K, ah.data_ptr(), al.data_ptr(), K, sc.data_ptr() + 4, bh.data_ptr(), bl.data_ptr(), Nn,
                  sc.data_ptr() + 12, Cm.data_ptr(), Nn, None, 1, 0, None, None, None, 0, st(main))
def cublas(n=40):
    for _ in range(n): torch.matmul(a, b)
def gather():
    _lib.call('ipavsr_gather_rows', host.data_ptr(), 4 * F, dev.data_ptr(), 4 * F, 4 * F, plan.pack.data_ptr(), None, plan.M + 1, st(cs))
def dma():
    with torch.cuda.stream(cs): dev.view(N, T, F).copy_(host, non_blocking=True)
def run(load, xfer, first='xfer'):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(main)
    if xfer and first == 'xfer':
        cs.wait_stream(main); xfer(); e2.record(cs)
    load()
    if xfer and first != 'xfer':
        xfer(); e2.record(cs)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), (e0.elapsed_time(e2) if xfer else 0.0)
for name, load in (('ours', ours), ('cublas', cublas)):
    for _ in range(2): run(load, None)
    print(name, 'alone            compute %.3f ms' % run(load, None)[0])
    for xn, xf in (('gather', gather), ('dma', dma)):
        for first in ('xfer', 'load'):
            r = run(load, xf, first)
            print(name, '+ %-6s (%s first) compute %.3f ms, transfer done at %.3f ms' % (xn, first, r[0], r[1]))
