"""fp16x3 vs tf32x3 GEMM timings at the bench shapes (CUDA events, L2 flushed)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipavsr_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')

def timeit(fn, reps=8):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps

for name, ta, tb, M, N, K in (('fc1 fwd', 0, 0, 20480, 2000, 1200), ('fc2 fwd', 0, 0, 20480, 1000, 2000), ('fc2 dgrad', 0, 1, 20480, 2000, 1000),
                              ('fc1 wgrad', 1, 0, 1200, 2000, 20480), ('fc2 wgrad', 1, 0, 2000, 1000, 20480), ('lstm proj', 0, 0, 20480, 1000, 152)):
    lda, ldb = (M if ta else K), (K if tb else N)
    A = torch.randn(K if ta else M, lda, device='cuda'); B = torch.randn(N if tb else K, ldb, device='cuda')
    Cm = torch.empty(M, N, device='cuda'); bias = torch.zeros(N, device='cuda')
    ah, al, bh, bl = (torch.empty_like(t, dtype=torch.float16) for t in (A, A, B, B))
    amax, exps = torch.zeros(4, device='cuda'), torch.zeros(4, dtype=torch.int32, device='cuda')
    _lib.call('ipavsr_f16_split', A.data_ptr(), lda, A.shape[0], A.shape[1], ah.data_ptr(), al.data_ptr(), lda, amax.data_ptr(), exps.data_ptr(), 0, st())
    _lib.call('ipavsr_f16_split', B.data_ptr(), ldb, B.shape[0], B.shape[1], bh.data_ptr(), bl.data_ptr(), ldb, amax.data_ptr() + 4, exps.data_ptr() + 4, 0, st())
    ms = timeit(lambda: _lib.call('ipavsr_gemm_f16x3', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, exps.data_ptr(), bh.data_ptr(), bl.data_ptr(), ldb,
                                  exps.data_ptr() + 4, Cm.data_ptr(), N, bias.data_ptr(), 1 if not ta else 0, 0, None, None, None, 0, st()))
    fl = 2.0 * M * N * K
    fh, fl32 = torch.empty_like(A), torch.empty_like(A); gh, gl = torch.empty_like(B), torch.empty_like(B)
    _lib.call('ipavsr_tf32_split_rna', A.data_ptr(), fh.data_ptr(), fl32.data_ptr(), A.numel(), st())
    _lib.call('ipavsr_tf32_split_rna', B.data_ptr(), gh.data_ptr(), gl.data_ptr(), B.numel(), st())
    ms3 = timeit(lambda: _lib.call('ipavsr_gemm_tf32x3_presplit', ta, tb, M, N, K, fh.data_ptr(), fl32.data_ptr(), lda, gh.data_ptr(), gl.data_ptr(), ldb,
                                   Cm.data_ptr(), N, bias.data_ptr(), 1 if not ta else 0, 0, None, None, st()))
    mss = timeit(lambda: _lib.call('ipavsr_f16_split', A.data_ptr(), lda, A.shape[0], A.shape[1], ah.data_ptr(), al.data_ptr(), lda, amax.data_ptr(), exps.data_ptr(), 0, st()))
    print('%-10s %6dx%5dx%6d  f16x3 %.3f ms %6.1f TF | tf32x3 %.3f ms %6.1f TF | split(A) %.3f ms %.0f GB/s' %
          (name, M, N, K, ms, fl / ms / 1e9, ms3, fl / ms3 / 1e9, mss, A.numel() * 12 / mss / 1e6), flush=True)
