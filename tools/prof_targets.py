"""One warm-up + one measured launch of each hot kernel at the bench shapes, for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:'delta_fwd_col|gemm_tc_kernel|lstm_.wd_persistent' \
        -o gpurun_out/prof_r01 python tools/prof_targets.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipavsr_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
which = set((sys.argv[1] if len(sys.argv) > 1 else 'delta,gemm,lstm').split(','))
T = 40
if 'delta' in which:
    N, F = 26214, 50
    x = torch.randn(N * T, 52, device='cuda'); y = torch.empty(N * T, 152, device='cuda')
    for exact in (0, 1):
        for _ in range(2):
            _lib.call('ipavsr_delta_fwd', x.data_ptr(), 52, y.data_ptr(), 152, N, T, F, 9, exact, st())
if 'gemm' in which:
    for (ta, tb, M, N, K) in ((0, 0, 20480, 2000, 1200), (1, 0, 1200, 2000, 20480)):
        lda, ldb = (M if ta else K), (K if tb else N)
        A = torch.randn(K if ta else M, lda, device='cuda'); B = torch.randn(N if tb else K, ldb, device='cuda')
        Cm = torch.empty(M, N, device='cuda'); bias = torch.zeros(N, device='cuda')
        ah, al, bh, bl = torch.empty_like(A), torch.empty_like(A), torch.empty_like(B), torch.empty_like(B)
        _lib.call('ipavsr_tf32_split_rna', A.data_ptr(), ah.data_ptr(), al.data_ptr(), A.numel(), st())
        _lib.call('ipavsr_tf32_split_rna', B.data_ptr(), bh.data_ptr(), bl.data_ptr(), B.numel(), st())
        for _ in range(2):
            _lib.call('ipavsr_gemm_tf32x3_presplit', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, bh.data_ptr(), bl.data_ptr(), ldb,
                      Cm.data_ptr(), N, bias.data_ptr(), 1 if not ta else 0, 0, None, None, st())
        for _ in range(2):
            _lib.call('ipavsr_gemm', 2, ta, tb, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, Cm.data_ptr(), N, bias.data_ptr(), 1 if not ta else 0, 0, None, 0, st())
if 'lstm' in which:
    N, H = 448, 250
    ldh = 252
    xw = torch.randn(N * T, 4 * H, device='cuda'); whid = torch.randn(H, 4 * H, device='cuda') * 0.05
    peep = torch.randn(3, H, device='cuda') * 0.1; z = torch.zeros(H, device='cuda')
    lens = torch.randint(12, T + 1, (N,), device='cuda')
    mask = (torch.arange(T, device='cuda')[None, :] < lens[:, None]).to(torch.uint8).contiguous()
    out, hprev = torch.zeros(N * T, ldh, device='cuda'), torch.zeros(N * T, ldh, device='cuda')
    gates, cell = torch.empty(N * T, 4 * H, device='cuda'), torch.empty(N * T, H, device='cuda')
    nbytes = lib.ipavsr_lstm_workspace_bytes(N, T, H); ws = torch.empty((nbytes + 3) // 4, device='cuda')
    dout, dg = torch.randn(N * T, ldh, device='cuda'), torch.empty(N * T, 4 * H, device='cuda')
    dpeep, dci, dhi = torch.zeros(3, H, device='cuda'), torch.zeros(H, device='cuda'), torch.zeros(H, device='cuda')
    for _ in range(2):
        _lib.call('ipavsr_lstm_fwd', xw.data_ptr(), whid.data_ptr(), peep.data_ptr(), z.data_ptr(), z.data_ptr(), mask.data_ptr(), out.data_ptr(),
                  gates.data_ptr(), cell.data_ptr(), hprev.data_ptr(), N, T, H, ldh, 0, 0, ws.data_ptr(), nbytes, st())
        _lib.call('ipavsr_lstm_bwd', dout.data_ptr(), whid.data_ptr(), peep.data_ptr(), z.data_ptr(), mask.data_ptr(), gates.data_ptr(), cell.data_ptr(),
                  dg.data_ptr(), dpeep.data_ptr(), dci.data_ptr(), dhi.data_ptr(), N, T, H, ldh, 0, 5.0, 0, 0, ws.data_ptr(), nbytes, st())
torch.cuda.synchronize()
print('done')
