"""Per-pair accounting of the persistent fp16 GEMM (gemm_f16p.cu; globaltimer stamps written when enabled).
    python tools/gemm_phases_p.py [ta tb M N K]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ipavsr_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
ta, tb, M, N, K = (int(v) for v in sys.argv[1:6]) if len(sys.argv) > 5 else (0, 0, 13325, 2000, 1200)
ACT = int(os.environ.get('ACT', '1' if not ta else '0'))
lda, ldb = (M if ta else K), (K if tb else N)
lda, ldb = (lda + 7) // 8 * 8, (ldb + 7) // 8 * 8
A = torch.randn(K if ta else M, lda, device='cuda'); B = torch.randn(N if tb else K, ldb, device='cuda')
Cm = torch.empty(M, N, device='cuda'); bias = torch.zeros(N, device='cuda')
ah, al, bh, bl = (torch.empty_like(t, dtype=torch.float16) for t in (A, A, B, B))
amax, exps = torch.zeros(4, device='cuda'), torch.zeros(4, dtype=torch.int32, device='cuda')
_lib.call('ipavsr_f16_split', A.data_ptr(), lda, A.shape[0], (M if ta else K), ah.data_ptr(), al.data_ptr(), lda, amax.data_ptr(), exps.data_ptr(), 0, st())
_lib.call('ipavsr_f16_split', B.data_ptr(), ldb, B.shape[0], (K if tb else N), bh.data_ptr(), bl.data_ptr(), ldb, amax.data_ptr() + 4, exps.data_ptr() + 4, 0, st())
run = lambda: _lib.call('ipavsr_gemm_f16x3', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, exps.data_ptr(), bh.data_ptr(), bl.data_ptr(), ldb,
                        exps.data_ptr() + 4, Cm.data_ptr(), N, bias.data_ptr(), ACT, 0, None, None, None, 0, st())
for _ in range(3):
    run()
buf = torch.zeros(8 * 4096, dtype=torch.int64, device='cuda')
lib.ipavsr_debug_gemm_timestamps(C.c_void_p(buf.data_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.ipavsr_debug_gemm_timestamps(None)
t = buf.cpu().numpy().reshape(-1, 8)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print('ta=%d tb=%d %dx%dx%d act=%d: %.3f ms, %d pairs, %d units' % (ta, tb, M, N, K, ACT, e0.elapsed_time(e1), len(t), t[:, 2].sum()))
print('pair start (us)    min %.1f  median %.1f  max %.1f' % tuple(np.percentile((t[:, 0] - t0) / 1e3, [0, 50, 100])))
print('setup              %.2f us' % ((t[:, 1] - t[:, 0]).mean() / 1e3))
print('units per pair     min %d  median %d  max %d' % tuple(np.percentile(t[:, 2], [0, 50, 100])))
print('pair lifetime (us) min %.1f  median %.1f  max %.1f' % tuple(np.percentile((t[:, 5] - t[:, 0]) / 1e3, [0, 50, 100])))
print('per unit: lifetime %.2f us | MMA thread waits: accumulator buffer %.2f us, operands %.2f us | epilogue busy %.2f us, waits for accumulators %.2f us'
      % (((t[:, 5] - t[:, 0]) / t[:, 2]).mean() / 1e3, (t[:, 3] / t[:, 2]).mean() / 1e3, (t[:, 4] / t[:, 2]).mean() / 1e3,
         (t[:, 6] / t[:, 2]).mean() / 1e3, (t[:, 7] / t[:, 2]).mean() / 1e3))
print('kernel span        %.1f us' % ((t[:, 5].max() - t0) / 1e3))
