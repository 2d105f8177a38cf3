import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipavsr_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(M, N, K, lda, reps=20):
    A = torch.randn(M, lda, device='cuda'); B = torch.randn(N, K, device='cuda')
    Cm = torch.zeros(M, N + 6, device='cuda')
    ah, al, bh, bl = (torch.empty_like(t, dtype=torch.float16) for t in (A, A, B, B))
    amax, exps = torch.zeros(4, device='cuda'), torch.zeros(4, dtype=torch.int32, device='cuda')
    _lib.call('ipavsr_f16_split', A.data_ptr(), lda, M, K, ah.data_ptr(), al.data_ptr(), lda, amax.data_ptr(), exps.data_ptr(), 0, st())
    _lib.call('ipavsr_f16_split', B.data_ptr(), K, N, K, bh.data_ptr(), bl.data_ptr(), K, amax.data_ptr() + 4, exps.data_ptr() + 4, 0, st())
    f = lambda acc: _lib.call('ipavsr_gemm_f16x3', 0, 1, M, N, K, ah.data_ptr(), al.data_ptr(), lda, exps.data_ptr(), bh.data_ptr(), bl.data_ptr(), K,
                          exps.data_ptr() + 4, Cm.data_ptr(), N + 6, None, 0, acc, None, None, None, 0, st())
    for acc in (0, 1):
        for _ in range(3): f(acc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): f(acc)
        e1.record(); torch.cuda.synchronize()
        print('M %d N %d K %d lda %d acc %d: %.1f us' % (M, N, K, lda, acc, e0.elapsed_time(e1) / reps * 1e3), flush=True)
run(512, 250, 1000, 1000); run(512, 250, 1000, 40000); run(26, 250, 1000, 40000); run(26, 250, 1000, 1000); run(512, 256, 1000, 1000); run(512, 256, 1024, 1024)
