"""Micro-benchmarks of the individual kernels through the C-ABI (CUDA events, L2 flushed between repetitions).
    python tools/bench_kernels.py [--only lstm,gemm,delta,pre] [--json out.json]
Reports time per launch, achieved algorithmic GB/s or TFLOP/s and the fraction of the measured peak
(MEASURED_PEAKS.json) — the SURVEY §8d config-4 sweep (delta / normalise over 1M frames) plus GEMM and LSTM shapes."""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ipavsr_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument('--only', default='delta,pre,feat,batch,gemm,lstm,opt')
ap.add_argument('--json', default=None)
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--frames', type=int, default=1048576)
args = ap.parse_args()
lib = _lib.load()
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
except Exception:
    pass
HBM = float(PEAKS.get('hbm_gbs', 6650.0))
TF = float(PEAKS.get('bf16_tflops', 1590.0))
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
results = []


def timeit(fn, reps=None, do_flush=True):
    reps = reps or args.reps
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def report(name, ms, bytes_=None, flops=None, extra=''):
    r = {'kernel': name, 'ms': ms}
    if bytes_ is not None:
        r.update(gbs=bytes_ / ms / 1e6, frac_hbm=bytes_ / ms / 1e6 / HBM)
    if flops is not None:
        r.update(tflops=flops / ms / 1e9, frac_bf16_peak=flops / ms / 1e9 / TF)
    results.append(r)
    print('%-58s %9.3f ms' % (name, ms) + ('  %8.1f GB/s (%4.1f%% of %d)' % (r['gbs'], 100 * r['frac_hbm'], HBM) if bytes_ else '') +
          ('  %8.1f TFLOP/s (%4.1f%% of bf16 %d)' % (r['tflops'], 100 * r['frac_bf16_peak'], TF) if flops else '') + extra, flush=True)


only = set(args.only.split(','))
T = 40
if 'delta' in only:
    N = args.frames // T
    for F in (30, 50, 90):
        ldx, ldy = (F + 3) // 4 * 4, (3 * F + 3) // 4 * 4
        x = torch.randn(N * T, ldx, device='cuda')
        y = torch.empty(N * T, ldy, device='cuda')
        for theta in (1, 4, 9):
            for exact in (1, 0):
                ms = timeit(lambda: _lib.call('ipavsr_delta_fwd', x.data_ptr(), ldx, y.data_ptr(), ldy, N, T, F, theta, exact, st()))
                report('delta_fwd F=%d theta=%d %s (%d frames)' % (F, theta, 'exact' if exact else 'fast', N * T), ms, bytes_=16.0 * F * N * T)
        g = torch.randn(N * T, ldy, device='cuda')
        ms = timeit(lambda: _lib.call('ipavsr_delta_bwd', g.data_ptr(), ldy, x.data_ptr(), ldx, N, T, F, 9, 0, st()))
        report('delta_bwd F=%d theta=9' % F, ms, bytes_=16.0 * F * N * T)
        del x, y, g
if 'pre' in only:
    frames, D = args.frames, 1200
    x = torch.randn(frames, D, device='cuda')
    y = torch.empty_like(x)
    ms = timeit(lambda: _lib.call('ipavsr_norm_samplewise', x.data_ptr(), D, y.data_ptr(), D, frames, D, st()))
    report('norm_samplewise D=1200 (%d frames)' % frames, ms, bytes_=8.0 * D * frames)
    mean, std = torch.empty(D, device='cuda'), torch.empty(D, device='cuda')
    scratch = torch.empty(3 * D, dtype=torch.float64, device='cuda')
    ms = timeit(lambda: _lib.call('ipavsr_norm_featurewise_stats', x.data_ptr(), D, mean.data_ptr(), std.data_ptr(), scratch.data_ptr(), frames, D, st()))
    report('norm_featurewise_stats D=1200 (two passes)', ms, bytes_=8.0 * D * frames)
    ms = timeit(lambda: _lib.call('ipavsr_norm_featurewise_apply', x.data_ptr(), D, mean.data_ptr(), std.data_ptr(), y.data_ptr(), D, frames, D, st()))
    report('norm_featurewise_apply D=1200', ms, bytes_=8.0 * D * frames)
    U = frames // T
    offs = torch.arange(0, U + 1, dtype=torch.int64, device='cuda') * T
    chunk = 65535
    def seq(fnname):
        for u0 in range(0, U, chunk):
            n = min(chunk, U - u0)
            _lib.call(fnname, x.data_ptr(), D, y.data_ptr(), D, offs.data_ptr() + 8 * u0, n, D, st())
    ms = timeit(lambda: seq('ipavsr_seq_mean_sub'))
    report('seq_mean_sub D=1200 T=40', ms, bytes_=8.0 * D * frames)
    ms = timeit(lambda: seq('ipavsr_diff_image'))
    report('diff_image D=1200 T=40', ms, bytes_=8.0 * D * frames)
    F = 30
    xf = torch.randn(frames, F, device='cuda')
    yf = torch.empty(frames, 3 * F, dtype=torch.float64, device='cuda')
    def fir():
        for u0 in range(0, U, chunk):
            n = min(chunk, U - u0)
            _lib.call('ipavsr_deltas_fir', xf.data_ptr(), F, yf.data_ptr(), 3 * F, offs.data_ptr() + 8 * u0, n, F, 9, T, st())
    ms = timeit(fir)
    report('deltas_fir F=30 w=9 (float64 out)', ms, bytes_=(4.0 + 24.0) * F * frames)
    del x, y, xf, yf
if 'feat' in only:
    # DCT projection on the zigzag basis vectors (4*D B read + 4*K B written per frame; 2*D*K flop per frame,
    # FP32-pipe bound), per-frame image reorder (8*D B/frame), force-align row gather (8*D B per output frame)
    from ipavsr_b200.utils import preprocessing as PP
    frames, D, K = args.frames // 2, 1200, 30
    x = torch.randn(frames, D, device='cuda')
    cols = torch.from_numpy(PP.zigzag_order(30, 40)[1:K + 1].astype(np.int32)).cuda()
    out = torch.empty(frames, K, device='cuda')
    for ldb, tag in ((32, 'cp.async path'), (K, 'generic path')):
        basis = torch.empty(D, ldb, device='cuda')
        _lib.call('ipavsr_dct_basis', basis.data_ptr(), ldb, cols.data_ptr(), D, K, st())
        ms = timeit(lambda: _lib.call('ipavsr_dct_project', x.data_ptr(), D, basis.data_ptr(), ldb, out.data_ptr(), K, frames, D, K, st()))
        report('dct_project D=%d K=%d %s (%d frames)' % (D, K, tag, frames), ms, bytes_=(4.0 * D + 4.0 * K) * frames, flops=2.0 * D * K * frames)
    y = torch.empty_like(x)
    ms = timeit(lambda: _lib.call('ipavsr_reorder', x.data_ptr(), D, y.data_ptr(), D, frames, 30, 40, 1, st()))
    report('reorder_data 30x40 f->c (%d frames)' % frames, ms, bytes_=8.0 * D * frames)
    U = frames // 32
    rng = np.random.default_rng(0)
    lin = rng.integers(12, 33, size=U).astype(np.int64)
    lin[-1] += frames - int(lin.sum()) if int(lin.sum()) < frames else 0
    lout = np.maximum(lin, rng.integers(12, 33, size=U))
    in_off = torch.from_numpy(np.concatenate([[0], np.cumsum(lin)]).astype(np.int64)).cuda()
    out_off_h = np.concatenate([[0], np.cumsum(lout)]).astype(np.int64)
    out_off = torch.from_numpy(out_off_h).cuda()
    rows_in, rows_out = int(lin.sum()), int(out_off_h[-1])
    xa = torch.randn(rows_in, D, device='cuda')
    ya = torch.empty(rows_out, D, device='cuda')
    ms = timeit(lambda: _lib.call('ipavsr_align_fill', xa.data_ptr(), D, ya.data_ptr(), D, in_off.data_ptr(), out_off.data_ptr(), None, U, D, rows_out, st()))
    report('align_fill D=%d (%d -> %d frames, %d utterances)' % (D, rows_in, rows_out, U), ms, bytes_=8.0 * D * rows_out)
    del x, y, xa, ya, out

if 'batch' in only:
    # padded (N, T, F) batch + mask gathered from a packed dataset resident in HBM
    rng = np.random.default_rng(0)
    U, Nb = 16384, 4096
    for F in (1200, 90):
        lens = rng.integers(20, T + 1, size=U)
        integ = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.int64)
        data = torch.randn(int(lens.sum()), F, device='cuda')
        d_int, d_len = torch.from_numpy(integ).cuda(), torch.from_numpy(lens.astype(np.int32)).cuda()
        idx = rng.permutation(U)[:Nb]
        d_idx = torch.from_numpy(idx.astype(np.int32)).cuda()
        X = torch.empty(Nb * T, F, device='cuda')
        mask = torch.empty(Nb, T, dtype=torch.uint8, device='cuda')
        ms = timeit(lambda: _lib.call('ipavsr_batch_gather', data.data_ptr(), F, d_int.data_ptr(), d_len.data_ptr(), d_idx.data_ptr(), None,
                                      X.data_ptr(), F, mask.data_ptr(), None, Nb, T, F, st()))
        report('batch_gather F=%d (%d utt x T=%d, mean len %.1f)' % (F, Nb, T, lens[idx].mean()), ms,
               bytes_=4.0 * F * (float(lens[idx].sum()) + Nb * T) + Nb * T)
        del data, X
if 'batch' in only:
    # frame argmax -> per-utterance vote -> confusion matrix, probabilities resident in HBM
    Nb, Cn = 4096, 26
    probs = torch.rand(Nb * T, Cn, device='cuda')
    lens_e = torch.randint(20, T + 1, (Nb,), device='cuda')
    mask_e = (torch.arange(T, device='cuda')[None, :] < lens_e[:, None]).to(torch.uint8).contiguous()
    y_e = torch.randint(0, Cn, (Nb,), device='cuda').to(torch.uint8)
    pred = torch.empty(Nb, dtype=torch.int32, device='cuda')
    conf, corr = torch.zeros(Cn, Cn, dtype=torch.int32, device='cuda'), torch.zeros(1, dtype=torch.int32, device='cuda')
    ms = timeit(lambda: _lib.call('ipavsr_vote_eval', probs.data_ptr(), Cn, mask_e.data_ptr(), y_e.data_ptr(), Nb, T, Cn, pred.data_ptr(),
                                  conf.data_ptr(), corr.data_ptr(), st()))
    report('vote_eval %d utt x T=%d x C=%d' % (Nb, T, Cn), ms, bytes_=4.0 * Cn * float(lens_e.sum().item()) + Nb * T)
if 'gemm' in only:
    shapes = [('fc1 fwd', 0, 0, 20480, 2000, 1200), ('fc2 fwd', 0, 0, 20480, 1000, 2000), ('fc3 fwd', 0, 0, 20480, 500, 1000),
              ('bottleneck fwd', 0, 0, 20480, 50, 500), ('fc2 dgrad', 0, 1, 20480, 2000, 1000), ('fc1 wgrad', 1, 0, 1200, 2000, 20480),
              ('fc2 wgrad', 1, 0, 2000, 1000, 20480), ('lstm proj', 0, 0, 20480, 1000, 150), ('lstm W_hid wgrad', 1, 0, 250, 1000, 20480),
              ('fc1 fwd 4096utt', 0, 0, 163840, 2000, 1200)]
    for name, ta, tb, M, N, K in shapes:
        lda, ldb = ((M if ta else K) + 3) // 4 * 4, ((K if tb else N) + 3) // 4 * 4
        A = torch.randn(K if ta else M, lda, device='cuda')
        B = torch.randn(N if tb else K, ldb, device='cuda')
        Cm = torch.empty(M, (N + 3) // 4 * 4, device='cuda')
        bias = torch.zeros(N, device='cuda')
        ah, al, bh, bl = (torch.empty_like(A), torch.empty_like(A), torch.empty_like(B), torch.empty_like(B))
        _lib.call('ipavsr_tf32_split_rna', A.data_ptr(), ah.data_ptr(), al.data_ptr(), A.numel(), st())
        _lib.call('ipavsr_tf32_split_rna', B.data_ptr(), bh.data_ptr(), bl.data_ptr(), B.numel(), st())
        fl = 2.0 * M * N * K
        for mode, label in ((0, 'fp32 CUDA-core'), (2, 'tf32 tcgen05'), (1, '3xTF32 tcgen05 presplit')):
            if mode == 1:
                fn = lambda: _lib.call('ipavsr_gemm_tf32x3_presplit', ta, tb, M, N, K, ah.data_ptr(), al.data_ptr(), lda, bh.data_ptr(), bl.data_ptr(), ldb,
                                       Cm.data_ptr(), Cm.shape[1], bias.data_ptr(), 0, 0, None, None, st())
            else:
                fn = lambda: _lib.call('ipavsr_gemm', mode, ta, tb, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, Cm.data_ptr(), Cm.shape[1],
                                       bias.data_ptr(), 0, 0, None, 0, st())
            if mode == 0 and M > 100000:
                continue
            ms = timeit(fn, reps=5)
            report('gemm %-18s %dx%dx%d %s' % (name, M, N, K, label), ms, flops=fl)
        del A, B, Cm, ah, al, bh, bl
if 'lstm' in only:
    shapes = ((26, 250, 150), (512, 250, 150), (26, 500, 150), (512, 500, 150), (4096, 250, 150))
    if os.environ.get('IPAVSR_BENCH_LSTM_N'):
        shapes = tuple((int(n), 250, 150) for n in os.environ['IPAVSR_BENCH_LSTM_N'].split(','))
    for N, H, I in shapes:
        ldh = (H + 3) // 4 * 4
        xw = torch.randn(N * T, 4 * H, device='cuda')
        whid = torch.randn(H, 4 * H, device='cuda') * 0.05
        peep = torch.randn(3, H, device='cuda') * 0.1
        ci, hi = torch.zeros(H, device='cuda'), torch.zeros(H, device='cuda')
        lens = torch.randint(12, T + 1, (N,), device='cuda')
        mask = (torch.arange(T, device='cuda')[None, :] < lens[:, None]).to(torch.uint8).contiguous()
        out, hprev = torch.zeros(N * T, ldh, device='cuda'), torch.zeros(N * T, ldh, device='cuda')
        gates, cell = torch.empty(N * T, 4 * H, device='cuda'), torch.empty(N * T, H, device='cuda')
        nbytes = lib.ipavsr_lstm_workspace_bytes(N, T, H)
        ws = torch.empty((nbytes + 3) // 4, device='cuda')
        dout, dg = torch.randn(N * T, ldh, device='cuda'), torch.empty(N * T, 4 * H, device='cuda')
        dpeep, dci, dhi = torch.zeros(3, H, device='cuda'), torch.zeros(H, device='cuda'), torch.zeros(H, device='cuda')
        for impl in (0, 1):
            if impl == 1 and N > 512:
                continue
            ms = timeit(lambda: _lib.call('ipavsr_lstm_fwd', xw.data_ptr(), whid.data_ptr(), peep.data_ptr(), ci.data_ptr(), hi.data_ptr(), mask.data_ptr(),
                                          out.data_ptr(), gates.data_ptr(), cell.data_ptr(), hprev.data_ptr(), N, T, H, ldh, 0, impl, ws.data_ptr(), nbytes, st()), reps=5)
            report('lstm_fwd N=%d H=%d impl=%d (train saves)' % (N, H, impl), ms, flops=2.0 * N * T * H * 4 * H, extra='  %.2f us/step' % (ms * 1e3 / T))
            ms = timeit(lambda: _lib.call('ipavsr_lstm_bwd', dout.data_ptr(), whid.data_ptr(), peep.data_ptr(), ci.data_ptr(), mask.data_ptr(), gates.data_ptr(),
                                          cell.data_ptr(), dg.data_ptr(), dpeep.data_ptr(), dci.data_ptr(), dhi.data_ptr(), N, T, H, ldh, 0, 5.0, 0, impl,
                                          ws.data_ptr(), nbytes, st()), reps=5)
            report('lstm_bwd N=%d H=%d impl=%d' % (N, H, impl), ms, flops=2.0 * N * T * H * 4 * H, extra='  %.2f us/step' % (ms * 1e3 / T))
        ms = timeit(lambda: _lib.call('ipavsr_lstm_fwd', xw.data_ptr(), whid.data_ptr(), peep.data_ptr(), ci.data_ptr(), hi.data_ptr(), mask.data_ptr(),
                                      out.data_ptr(), None, None, None, N, T, H, ldh, 0, 0, ws.data_ptr(), nbytes, st()), reps=5)
        report('lstm_fwd N=%d H=%d impl=0 (inference)' % (N, H), ms, flops=2.0 * N * T * H * 4 * H, extra='  %.2f us/step' % (ms * 1e3 / T))
        if lib.ipavsr_lstm_fwd_f16_supported(N, T, H, 4 * H):
            wh, wl = torch.empty(H, 4 * H, dtype=torch.float16, device='cuda'), torch.empty(H, 4 * H, dtype=torch.float16, device='cuda')
            sc = torch.zeros(2, device='cuda')
            _lib.call('ipavsr_f16_split', whid.data_ptr(), 4 * H, H, 4 * H, wh.data_ptr(), wl.data_ptr(), 4 * H, sc.data_ptr(), sc.data_ptr() + 4, 0, st())
            ms = timeit(lambda: _lib.call('ipavsr_lstm_fwd_f16', xw.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, 4 * H, peep.data_ptr(), ci.data_ptr(),
                                          hi.data_ptr(), mask.data_ptr(), out.data_ptr(), gates.data_ptr(), cell.data_ptr(), hprev.data_ptr(), N, T, H, ldh, 0, st()), reps=5)
            report('lstm_fwd_f16 N=%d H=%d tcgen05 (train saves)' % (N, H), ms, flops=2.0 * N * T * H * 4 * H, extra='  %.2f us/step' % (ms * 1e3 / T))
            ms = timeit(lambda: _lib.call('ipavsr_lstm_bwd_f16', dout.data_ptr(), whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, 4 * H, peep.data_ptr(),
                                          ci.data_ptr(), mask.data_ptr(), gates.data_ptr(), cell.data_ptr(), dg.data_ptr(), dpeep.data_ptr(), dci.data_ptr(), dhi.data_ptr(),
                                          N, T, H, ldh, 0, 5.0, 0, None, None, None, None, ws.data_ptr(), nbytes, st()), reps=5)
            report('lstm_bwd_f16 N=%d H=%d tcgen05' % (N, H), ms, flops=2.0 * N * T * H * 4 * H, extra='  %.2f us/step' % (ms * 1e3 / T))
        if H > 256 and lib.ipavsr_lstm_steps_supported(N, T, H, 4 * H):
            # wide layers: one tensor-core GEMM + one cell kernel per step (csrc/lstm_steps_tc.cu), unsorted lengths
            wh, wl = torch.empty(H, 4 * H, dtype=torch.float16, device='cuda'), torch.empty(H, 4 * H, dtype=torch.float16, device='cuda')
            sc = torch.zeros(2, device='cuda')
            _lib.call('ipavsr_f16_split', whid.data_ptr(), 4 * H, H, 4 * H, wh.data_ptr(), wl.data_ptr(), 4 * H, sc.data_ptr(), sc.data_ptr() + 4, 0, st())
            sb = lib.ipavsr_lstm_steps_workspace_bytes(N, T, H)
            sws = torch.empty((sb + 3) // 4, device='cuda')
            dgh, dgl = torch.empty(N * T, 4 * H, dtype=torch.float16, device='cuda'), torch.empty(N * T, 4 * H, dtype=torch.float16, device='cuda')
            dge = torch.zeros(2, dtype=torch.int32, device='cuda')
            ms = timeit(lambda: _lib.call('ipavsr_lstm_fwd_f16_steps', xw.data_ptr(), whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, 4 * H,
                                          peep.data_ptr(), ci.data_ptr(), hi.data_ptr(), mask.data_ptr(), out.data_ptr(), gates.data_ptr(), cell.data_ptr(),
                                          hprev.data_ptr(), N, T, H, ldh, 0, None, sws.data_ptr(), sb, st()), reps=5)
            report('lstm_fwd_f16_steps N=%d H=%d tcgen05 GEMM per step (train saves)' % (N, H), ms, flops=2.0 * N * T * H * 4 * H, extra='  %.2f us/step' % (ms * 1e3 / T))
            ms = timeit(lambda: _lib.call('ipavsr_lstm_bwd_f16_steps', dout.data_ptr(), whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, 4 * H,
                                          peep.data_ptr(), ci.data_ptr(), mask.data_ptr(), gates.data_ptr(), cell.data_ptr(), dg.data_ptr(), dpeep.data_ptr(),
                                          dci.data_ptr(), dhi.data_ptr(), N, T, H, ldh, 0, 5.0, 0, dgh.data_ptr(), dgl.data_ptr(), dge.data_ptr(), None,
                                          sws.data_ptr(), sb, st()), reps=5)
            report('lstm_bwd_f16_steps N=%d H=%d tcgen05 GEMM per step' % (N, H), ms, flops=2.0 * N * T * H * 4 * H, extra='  %.2f us/step' % (ms * 1e3 / T))
        del xw, out, hprev, gates, cell, dout, dg, ws
if 'opt' in only:
    n = 24 * 1024 * 1024
    p, g, m, v = (torch.randn(n, device='cuda') for _ in range(4))
    v.abs_()
    ms = timeit(lambda: _lib.call('ipavsr_optim_step', 0, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, None, None, 1.0, 0.9, 0.999, 1e-8, 1.0, st()))
    report('optim_step adam %d params' % n, ms, bytes_=28.0 * n)
if args.json:
    json.dump({'peaks': {'hbm_gbs': HBM, 'bf16_tflops': TF}, 'results': results}, open(args.json, 'w'), indent=1)
