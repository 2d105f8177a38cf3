"""One warm-up + one measured launch of each hot kernel at the bench shapes, for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on \
        -k regex:'delta_stream|gemm_tc_kernel|lstm_(fwd|bwd)_tc' -o gpurun_out/prof_r01 python tools/prof_targets.py"""
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
    x = torch.randn(N * T, 56, device='cuda'); y = torch.empty(N * T, 152, device='cuda'); gx = torch.empty(N * T, 56, device='cuda')
    for exact in (0, 1):
        for _ in range(2):
            _lib.call('ipavsr_delta_fwd', x.data_ptr(), 56, y.data_ptr(), 152, N, T, F, 9, exact, st())
    for _ in range(2):
        _lib.call('ipavsr_delta_bwd', y.data_ptr(), 152, gx.data_ptr(), 56, N, T, F, 9, 0, st())
if 'gemm' in which:
    R = int(os.environ.get('IPAVSR_PROF_ROWS', '38400'))          # batch * T rows of the bench workload (960 x 40)
    for (ta, tb, M, N, K) in ((0, 0, R, 2000, 1200), (1, 0, 1200, 2000, R)):
        lda, ldb = (M if ta else K), (K if tb else N)
        A = torch.randn(K if ta else M, lda, device='cuda'); B = torch.randn(N if tb else K, ldb, device='cuda')
        Cm = torch.empty(M, N, device='cuda'); bias = torch.zeros(N, device='cuda')
        ah, al, bh, bl = (torch.empty_like(t, dtype=torch.float16) for t in (A, A, B, B))
        sc = torch.zeros(4, device='cuda')
        _lib.call('ipavsr_f16_split', A.data_ptr(), lda, A.shape[0], A.shape[1], ah.data_ptr(), al.data_ptr(), lda, sc.data_ptr(), sc.data_ptr() + 4, 0, st())
        _lib.call('ipavsr_f16_split', B.data_ptr(), ldb, B.shape[0], B.shape[1], bh.data_ptr(), bl.data_ptr(), ldb, sc.data_ptr() + 8, sc.data_ptr() + 12, 0, st())
        for _ in range(2):
            _lib.call('ipavsr_gemm_f16x3', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, sc.data_ptr() + 4, bh.data_ptr(), bl.data_ptr(), ldb,
                      sc.data_ptr() + 12, Cm.data_ptr(), N, bias.data_ptr(), 1 if not ta else 0, 0, None, None, None, 0, st())
if 'lstm' in which:
    N, H = 480, 250
    ldh = 256
    xw = torch.randn(N * T, 4 * H, device='cuda'); whid = torch.randn(H, 4 * H, device='cuda') * 0.05
    peep = torch.randn(3, H, device='cuda') * 0.1; z = torch.zeros(H, device='cuda')
    lens = torch.randint(12, T + 1, (N,), device='cuda')
    mask = (torch.arange(T, device='cuda')[None, :] < lens[:, None]).to(torch.uint8).contiguous()
    out, hprev = torch.zeros(N * T, ldh, device='cuda'), torch.zeros(N * T, ldh, device='cuda')
    gates, cell = torch.empty(N * T, 4 * H, device='cuda'), torch.empty(N * T, H, device='cuda')
    nbytes = lib.ipavsr_lstm_workspace_bytes(N, T, H); ws = torch.empty((nbytes + 3) // 4, device='cuda')
    dout, dg = torch.randn(N * T, ldh, device='cuda'), torch.empty(N * T, 4 * H, device='cuda')
    dpeep, dci, dhi = torch.zeros(3, H, device='cuda'), torch.zeros(H, device='cuda'), torch.zeros(H, device='cuda')
    wh, wl = torch.empty(H, 4 * H, dtype=torch.float16, device='cuda'), torch.empty(H, 4 * H, dtype=torch.float16, device='cuda')
    sc = torch.zeros(2, device='cuda')
    _lib.call('ipavsr_f16_split', whid.data_ptr(), 4 * H, H, 4 * H, wh.data_ptr(), wl.data_ptr(), 4 * H, sc.data_ptr(), sc.data_ptr() + 4, 0, st())
    for _ in range(2):
        _lib.call('ipavsr_lstm_fwd_f16', xw.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, 4 * H, peep.data_ptr(), z.data_ptr(), z.data_ptr(),
                  mask.data_ptr(), out.data_ptr(), gates.data_ptr(), cell.data_ptr(), hprev.data_ptr(), N, T, H, ldh, 0, st())
        _lib.call('ipavsr_lstm_bwd_f16', dout.data_ptr(), whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, 4 * H, peep.data_ptr(), z.data_ptr(),
                  mask.data_ptr(), gates.data_ptr(), cell.data_ptr(), dg.data_ptr(), dpeep.data_ptr(), dci.data_ptr(), dhi.data_ptr(), N, T, H, ldh, 0, 5.0, 0,
                  None, None, None, None, ws.data_ptr(), nbytes, st())
torch.cuda.synchronize()
print('done')
