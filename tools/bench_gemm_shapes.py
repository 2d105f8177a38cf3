"""f16x3 GEMM timings at the shapes of the default 960-utterance step (CUDA events, L2 flushed): finds shapes whose tile
count / k-depth sits badly on the 148 SMs.   python tools/bench_gemm_shapes.py [rows]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipavsr_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
R = int(sys.argv[1]) if len(sys.argv) > 1 else 38400


def timeit(fn, reps=6):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


ld8 = lambda n: (n + 7) // 8 * 8
shapes = [('fc1 fwd', 0, 0, R, 2000, 1200, 1), ('fc2 fwd', 0, 0, R, 1000, 2000, 1), ('fc3 fwd', 0, 0, R, 500, 1000, 1), ('bneck fwd', 0, 0, R, 50, 500, 0),
          ('lstm proj', 0, 0, R, 1000, 150, 0), ('blstm proj K=750', 0, 0, R, 1000, 750, 0), ('blstm proj K=768', 0, 0, R, 1000, 768, 0),
          ('fc2 dgrad', 0, 1, R, 2000, 1000, 0), ('fc3 dgrad', 0, 1, R, 1000, 500, 0), ('blstm dgrad', 0, 1, R, 750, 1000, 0), ('lstm dgrad', 0, 1, R, 150, 1000, 0),
          ('fc1 wgrad', 1, 0, 1200, 2000, R, 0), ('fc2 wgrad', 1, 0, 2000, 1000, R, 0), ('fc3 wgrad', 1, 0, 1000, 500, R, 0), ('blstm wgrad', 1, 0, 750, 1000, R, 0),
          ('whid wgrad', 1, 0, 250, 1000, R, 0), ('lstm wgrad', 1, 0, 150, 1000, R, 0)]
for name, ta, tb, M, N, K, act in shapes:
    lda, ldb, ldc = ld8(M if ta else K), ld8(K if tb else N), ld8(N)
    A = torch.randn(K if ta else M, lda, device='cuda'); B = torch.randn(N if tb else K, ldb, device='cuda')
    Cm = torch.empty(M, ldc, device='cuda'); bias = torch.zeros(N, device='cuda')
    ah, al, bh, bl = (torch.empty_like(t, dtype=torch.float16) for t in (A, A, B, B))
    amax, exps = torch.zeros(4, device='cuda'), torch.zeros(4, dtype=torch.int32, device='cuda')
    _lib.call('ipavsr_f16_split', A.data_ptr(), lda, A.shape[0], (M if ta else K), ah.data_ptr(), al.data_ptr(), lda, amax.data_ptr(), exps.data_ptr(), 0, st())
    _lib.call('ipavsr_f16_split', B.data_ptr(), ldb, B.shape[0], (K if tb else N), bh.data_ptr(), bl.data_ptr(), ldb, amax.data_ptr() + 4, exps.data_ptr() + 4, 0, st())
    ms = timeit(lambda: _lib.call('ipavsr_gemm_f16x3', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, exps.data_ptr(), bh.data_ptr(), bl.data_ptr(), ldb,
                                  exps.data_ptr() + 4, Cm.data_ptr(), ldc, bias.data_ptr(), act, 0, None, None, None, 0, st()))
    print('%-18s ta=%d tb=%d %6dx%5dx%6d  %.3f ms %6.1f TFLOP/s' % (name, ta, tb, M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
    del A, B, Cm, ah, al, bh, bl
