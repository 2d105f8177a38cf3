"""Per-CTA phase timing of the tensor-core GEMM (globaltimer stamps written by the kernel when enabled).
    python tools/gemm_phases.py [f16|tf32x3] [ta tb M N K]"""
import ctypes as C, os, sys
os.environ.setdefault('IPAVSR_GEMM_PERSIST', '0')     # this tool reads the tile-per-pair kernel's stamps; the persistent
                                                      # kernel (gemm_f16p.cu) has its own accounting: tools/gemm_phases_p.py
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ipavsr_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
kind = sys.argv[1] if len(sys.argv) > 1 else 'f16'
ta, tb, M, N, K = (int(v) for v in sys.argv[2:7]) if len(sys.argv) > 6 else (0, 0, 20480, 2000, 1200)
ACT = int(os.environ.get('ACT', '1' if not ta else '0'))
USE_BIAS = os.environ.get('BIAS', '1') == '1'
lda, ldb = (M if ta else K), (K if tb else N)
A = torch.randn(K if ta else M, lda, device='cuda'); B = torch.randn(N if tb else K, ldb, device='cuda')
Cm = torch.empty(M, N, device='cuda'); bias = torch.zeros(N, device='cuda')
if kind == 'f16':
    ah, al, bh, bl = (torch.empty_like(t, dtype=torch.float16) for t in (A, A, B, B))
    amax, exps = torch.zeros(4, device='cuda'), torch.zeros(4, dtype=torch.int32, device='cuda')
    _lib.call('ipavsr_f16_split', A.data_ptr(), lda, A.shape[0], A.shape[1], ah.data_ptr(), al.data_ptr(), lda, amax.data_ptr(), exps.data_ptr(), 0, st())
    _lib.call('ipavsr_f16_split', B.data_ptr(), ldb, B.shape[0], B.shape[1], bh.data_ptr(), bl.data_ptr(), ldb, amax.data_ptr() + 4, exps.data_ptr() + 4, 0, st())
    run = lambda: _lib.call('ipavsr_gemm_f16x3', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, exps.data_ptr(), bh.data_ptr(), bl.data_ptr(), ldb,
                            exps.data_ptr() + 4, Cm.data_ptr(), N, bias.data_ptr() if USE_BIAS else None, ACT, 0, None, None, None, 0, st())
else:
    ah, al, bh, bl = torch.empty_like(A), torch.empty_like(A), torch.empty_like(B), torch.empty_like(B)
    _lib.call('ipavsr_tf32_split_rna', A.data_ptr(), ah.data_ptr(), al.data_ptr(), A.numel(), st())
    _lib.call('ipavsr_tf32_split_rna', B.data_ptr(), bh.data_ptr(), bl.data_ptr(), B.numel(), st())
    run = lambda: _lib.call('ipavsr_gemm_tf32x3_presplit', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, bh.data_ptr(), bl.data_ptr(), ldb,
                            Cm.data_ptr(), N, bias.data_ptr() if USE_BIAS else None, ACT, 0, None, None, st())
for _ in range(3):
    run()
buf = torch.zeros(8 * 65536, dtype=torch.int64, device='cuda')
lib.ipavsr_debug_gemm_timestamps(C.c_void_p(buf.data_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.ipavsr_debug_gemm_timestamps(None)
t = buf.cpu().numpy().reshape(-1, 8)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print('%s ta=%d tb=%d %dx%dx%d act=%d bias=%d: %.3f ms, %d CTAs' % (kind, ta, tb, M, N, K, ACT, USE_BIAS, e0.elapsed_time(e1), len(t)))
lead = t[t[:, 3] > 0]          # CTAs that issued MMAs
d = lambda a, b, rows=t: (rows[:, b] - rows[:, a]) / 1e3
print('setup            %6.2f us (mean)  %6.2f max' % (d(0, 1).mean(), d(0, 1).max()))
print('first stage wait %6.2f us' % d(1, 2, lead).mean())
print('mainloop issue   %6.2f us (first stage -> last MMA issued)' % d(2, 3, lead).mean())
print('start -> acc rdy %6.2f us' % d(0, 4).mean())
print('epilogue         %6.2f us (mean)  %6.2f max' % (d(4, 5).mean(), d(4, 5).max()))
print('CTA lifetime     %6.2f us' % d(0, 5).mean())
print('kernel span      %6.2f us (first start -> last epilogue end)' % ((t[:, 5].max() - t0) / 1e3))
starts = np.sort(t[:, 0] - t0) / 1e3
print('CTA start times (us) deciles:', np.round(starts[(np.linspace(0, 1, 11) * (len(starts) - 1)).astype(int)], 1))
