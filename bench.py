#!/usr/bin/env python
"""Benchmark of the AdeNet hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch B] [--mode f16x3|tf32x3|fp32|tf32]
                    [--no-extras] [--no-cpu-baseline]

Headline workload (config.workload; BASELINE.json config 3): the AdeNet-v2 late-fusion *trimodal* network
(`modelzoo.adenet_3stream`: raw 1200-px ROI + diff-image 1200-px + DCT 90, each through a DBNF encoder 2000-1000-500-50 ->
DeltaLayer(theta=9) -> masked LSTM-250 with peepholes, concat fusion, BLSTM-250 aggregate, per-frame softmax over 26
classes), one training step = forward + temporal_softmax_loss + full backward + Adam, 512 utterances per GPU, T=40 padded
frames, lengths U{12..40}, synthetic data, random-init weights.  Per-GPU batch is fixed (weak scaling); with N>1 one
process per GPU (torchrun) shards utterances and all-reduces the flat gradient arena over NCCL.

One JSON line on stdout (rank 0):
  value        utterances/s with the three padded streams already resident in HBM (device-timed, max over ranks);
  e2e          the same step through the public `function(...)` callable from pinned HOST memory: every step uploads the
               valid frames of the raw stream (ragged copy-engine transfer), computes the diff-image and DCT(+deltas)
               streams on the device from it (ipavsr_b200.derived) and reads the loss back;
  roofline     the dominant kernel (encoder fc1 GEMM on the packed rows of the batch) timed live, vs the measured bf16 peak;
  rooflines    BASELINE config 4: the streaming kernels over 1 048 576 frames vs the measured HBM bandwidth;
  inference_4stream   BASELINE config 5: adenet_4stream, 4096 utterances sharded over the ranks, no collective;
  configs      BASELINE configs 1-3 at the reference's own batch sizes and at 512 (rank 0, single GPU only);
  cpu_baseline the NumPy restatement of the reference path on the host cores (bounded sample).
`--impl reference` times that CPU restatement alone (Theano/Lasagne cannot be installed here).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_FRAMES, THETA, H_LSTM, N_CLASSES = 40, 9, 250, 26
STREAM_DIMS = (1200, 1200, 90)
IMAGE_SHAPE, DCT_COEFF = (30, 40), 30
ENC = (2000, 1000, 500, 50)
ENC_ACTS = ('sigmoid', 'sigmoid', 'sigmoid', 'linear')


def make_ae(rng, D, acts=ENC_ACTS):
    from ipavsr_b200 import nonlinearities as nl
    s = (D,) + ENC
    W = [rng.normal(0, 1.0 / np.sqrt(s[i]), (s[i], s[i + 1])).astype('float32') for i in range(4)]
    b = [rng.normal(0, 0.1, (s[i + 1],)).astype('float32') for i in range(4)]
    return W, b, list(ENC), [nl.select_nonlinearity(a) for a in acts]


class _DBN(object):
    """The legacy encoder object form (get_all_layers()[1..4] carry .W/.b, modelzoo/deltanet.py:63-73)."""

    class _L(object):
        def __init__(self, W, b):
            self.W, self.b = W, b

    def __init__(self, ae):
        self._layers = [None] + [_DBN._L(w, b) for w, b in zip(ae[0], ae[1])]

    def get_all_layers(self):
        return self._layers


def build_network(seed=1234):
    from ipavsr_b200 import modelzoo, init
    from ipavsr_b200.function import tensor as T
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    aes = [make_ae(rng, D) for D in STREAM_DIMS]
    v = [T.tensor3('s%d' % (i + 1)) for i in range(3)]
    mask = T.matrix('mask', dtype='uint8')
    window = T.iscalar('theta')
    net, fuse = modelzoo.adenet_3stream.create_model(
        aes[0], aes[1], aes[2], (None, None, STREAM_DIMS[0]), v[0], (None, None, STREAM_DIMS[1]), v[1],
        (None, None, STREAM_DIMS[2]), v[2], (None, None), mask, H_LSTM, window, N_CLASSES, 'concat',
        init.Orthogonal(), True)
    return net, v, mask, window


def synth_batch(n, seed, dims=STREAM_DIMS, lo=12, classes=N_CLASSES, frame_targets=True):
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, T_FRAMES + 1, size=n)
    lens[0] = T_FRAMES
    mask = (np.arange(T_FRAMES)[None, :] < lens[:, None]).astype('uint8')
    xs = []
    for D in dims:
        x = rng.standard_normal((n, T_FRAMES, D), dtype=np.float32)
        x *= mask[:, :, None]
        xs.append(x)
    y = rng.integers(0, classes, size=(n, 1))
    y = (np.repeat(y, T_FRAMES, 1) if frame_targets else y[:, 0]).astype('int32')
    return xs, mask, y


def flops_per_utt_train():
    """Algorithmic FLOPs of one training step per utterance of T padded frames (fwd + dgrad + wgrad; first-layer dgrad not
    needed) — the reference's arithmetic, which computes every padded frame."""
    T = T_FRAMES
    f = 0.0
    for D in STREAM_DIMS:
        dims = (D,) + ENC
        fwd = sum(2.0 * dims[i] * dims[i + 1] for i in range(4))
        f += T * (3 * fwd - 2.0 * dims[0] * dims[1])
        f += T * 3 * 2.0 * 150 * 4 * H_LSTM + T * 3 * 2.0 * H_LSTM * 4 * H_LSTM
    f += 2 * (T * 3 * 2.0 * 3 * H_LSTM * 4 * H_LSTM + T * 3 * 2.0 * H_LSTM * 4 * H_LSTM)
    f += T * 3 * 2.0 * H_LSTM * N_CLASSES
    return f


class ClockSampler(object):
    """Samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.samples, self.stop, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_reference_step_rate(n_utt, steps, warmup, seed=99):
    """The reference path on the CPU: NumPy float32 oracle (BLAS dots; vectorised delta), forward + loss + backward
    + Adam on `n_utt` utterances per step."""
    from ipavsr_b200 import layers as L
    from oracle.net import OracleNet
    from oracle import ops
    net, v, mask_var, window = build_network()
    params = L.get_all_params(net, trainable=True)
    xs, mask, y = synth_batch(n_utt, seed)
    feed = {'s1_im': xs[0], 's2_im': xs[1], 's3_im': xs[2], 'mask': mask}
    o = OracleNet(net, np.float32)
    st = {'t': np.float32(0), 'm': [np.zeros(p.shape, 'f') for p in params], 'v': [np.zeros(p.shape, 'f') for p in params]}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, grads = o.loss_and_grads(feed, THETA, y, mask, 'temporal_softmax', deterministic=False)
        cur = [p._host for p in params]
        ops.adam_step(cur, grads, st, [1e-3] * len(cur))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_utt / (sum(times) / len(times)), sum(times) / len(times)


# ---------------------------------------------------------------------------------------------------------
# timing helpers (device side)
# ---------------------------------------------------------------------------------------------------------
def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def time_kernel(fn, flush, reps=8, warm=2):
    """Average device time of one launch: CUDA events on the launching stream, L2 flushed before every repetition."""
    import torch
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def streaming_rooflines(hbm_peak, peak_source, frames=1048576):
    """BASELINE config 4 / SURVEY 8(d): DeltaLayer and utils/preprocessing kernels over `frames` frames; achieved =
    algorithmic bytes / average launch time, against the measured HBM copy bandwidth."""
    import torch
    from ipavsr_b200 import _lib
    T = T_FRAMES
    N = frames // T
    rows = N * T
    flush = torch.empty(192 * 1024 * 1024 // 4, device='cuda')
    out = []

    def rec(name, ms, nbytes, note=None):
        gbs = nbytes / ms / 1e6
        r = {'kernel': name, 'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
             'ms_per_launch': ms, 'algorithmic_bytes': nbytes, 'frames': rows, 'peak_source': peak_source}
        if note:
            r['note'] = note
        out.append(r)

    for F in (50, 90, 30):
        ldx, ldy = (F + 7) // 8 * 8, (3 * F + 7) // 8 * 8
        x = torch.randn(rows, ldx, device='cuda')
        y = torch.empty(rows, ldy, device='cuda')
        for theta in ((1, 4, 9) if F == 50 else (9,)):
            for exact in (1, 0):
                ms = time_kernel(lambda: _lib.call('ipavsr_delta_fwd', x.data_ptr(), ldx, y.data_ptr(), ldy, N, T, F, theta,
                                                   exact, _stream()), flush)
                rec('delta_fwd F=%d theta=%d %s' % (F, theta, 'exact (engine default: bit-exact float64 chain of '
                                                    'utils/signal.py:19-21)' if exact else 'fast (float32 chain, 1e-6 rel)'),
                    ms, 16.0 * F * rows)
        g = torch.randn(rows, ldy, device='cuda')
        ms = time_kernel(lambda: _lib.call('ipavsr_delta_bwd', g.data_ptr(), ldy, x.data_ptr(), ldx, N, T, F, 9, 0, _stream()),
                         flush)
        rec('delta_bwd F=%d theta=9' % F, ms, 16.0 * F * rows)
        del x, y, g
    D = 1200
    x = torch.randn(rows, D, device='cuda')
    y = torch.empty_like(x)
    ms = time_kernel(lambda: _lib.call('ipavsr_norm_samplewise', x.data_ptr(), D, y.data_ptr(), D, rows, D, _stream()), flush)
    rec('norm_samplewise D=1200 (normalize_input)', ms, 8.0 * D * rows)
    mean, std = torch.zeros(D, device='cuda'), torch.ones(D, device='cuda')
    ms = time_kernel(lambda: _lib.call('ipavsr_norm_featurewise_apply', x.data_ptr(), D, mean.data_ptr(), std.data_ptr(),
                                       y.data_ptr(), D, rows, D, _stream()), flush)
    rec('featurewise_apply D=1200', ms, 8.0 * D * rows)
    scratch = torch.empty(3 * D, dtype=torch.float64, device='cuda')
    ms = time_kernel(lambda: _lib.call('ipavsr_norm_featurewise_stats', x.data_ptr(), D, mean.data_ptr(), std.data_ptr(),
                                       scratch.data_ptr(), rows, D, _stream()), flush)
    rec('featurewise_stats D=1200 (two passes over x)', ms, 8.0 * D * rows)
    offs = torch.arange(0, N + 1, dtype=torch.int64, device='cuda') * T

    def chunked(name):
        for u0 in range(0, N, 65535):
            _lib.call(name, x.data_ptr(), D, y.data_ptr(), D, offs.data_ptr() + 8 * u0, min(65535, N - u0), D, _stream())
    ms = time_kernel(lambda: chunked('ipavsr_seq_mean_sub'), flush)
    rec('seq_mean_sub D=1200 T=40', ms, 8.0 * D * rows)
    ms = time_kernel(lambda: chunked('ipavsr_diff_image'), flush)
    rec('diff_image D=1200 T=40', ms, 8.0 * D * rows)
    # ragged pack of the valid frames (the engine's packed execution), lengths U{12..40}
    rng = np.random.default_rng(0)
    lens = rng.integers(12, T + 1, size=N)
    from ipavsr_b200.engine import _PackPlan
    plan = _PackPlan(lens, T)
    pin = torch.empty(sum(len(a) for _, a in plan.tables), dtype=torch.int32).pin_memory()
    plan.upload(torch.device('cuda', torch.cuda.current_device()), pin)
    ms = time_kernel(lambda: _lib.call('ipavsr_gather_rows', x.data_ptr(), 4 * D, y.data_ptr(), 4 * D, 4 * D,
                                       plan.pack.data_ptr(), None, plan.M + 1, _stream()), flush)
    rec('gather_rows pack D=1200 (valid frames of %d padded)' % rows, ms, 8.0 * D * plan.M)
    del x, y
    F = 30
    xf = torch.randn(rows, F, device='cuda')
    y64 = torch.empty(rows, 3 * F, dtype=torch.float64, device='cuda')
    y32 = torch.empty(rows, 3 * F, dtype=torch.float32, device='cuda')
    ms = time_kernel(lambda: _lib.call('ipavsr_deltas_fir', xf.data_ptr(), F, y64.data_ptr(), 3 * F, offs.data_ptr(), N, F, 9, T,
                                       _stream()), flush)
    rec('deltas_fir F=30 w=9, float64 output (the reference array)', ms, 28.0 * F * rows,
        note='SURVEY 8d counts 16F B/frame for float32 in and out; the reference returns float64: 4F read + 24F written')
    ms = time_kernel(lambda: _lib.call('ipavsr_deltas_fir_f32', xf.data_ptr(), F, y32.data_ptr(), 3 * F, offs.data_ptr(), N, F, 9,
                                       T, _stream()), flush)
    rec('deltas_fir_f32 F=30 w=9, float32 output (what the runners feed the network)', ms, 16.0 * F * rows)
    return out


def build_config_net(name, seed):
    """Networks of BASELINE configs 1-3 at the reference's layer sizes.  Returns (net, input vars in call order, mask var,
    window var, stream dims, frame-level?, classes, dropout?)."""
    from ipavsr_b200 import modelzoo, init
    from ipavsr_b200.function import tensor as T
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    m, w = T.matrix('mask', dtype='uint8'), T.iscalar('theta')
    sh = lambda d: (None, None, d)
    if name == 'deltanet':                      # config 1: AVLetters unimodal, modelzoo/deltanet.py:59 (rectify encoder)
        v = [T.tensor3('x')]
        net = modelzoo.deltanet.create_model(_DBN(make_ae(rng, 1200)), sh(1200), v[0], (None, None), m, 250, w, 26)
        return net, v, m, w, [1200], False, 26
    if name == 'adenet_v1':                     # config 2: OuluVS bimodal early fusion, modelzoo/adenet_v1.py:47
        v = [T.tensor3('x'), T.tensor3('dct')]
        net, _ = modelzoo.adenet_v1.create_model(_DBN(make_ae(rng, 1144)), sh(1144), v[0], (None, None), m, sh(90), v[1],
                                                 250, w, 10)
        return net, v, m, w, [1144, 90], False, 10
    if name == 'adenet_v2':                     # config 2: late fusion concat, modelzoo/adenet_v2.py:12
        v = [T.tensor3('x'), T.tensor3('dct')]
        net, _ = modelzoo.adenet_v2.create_model(make_ae(rng, 1144), sh(1144), v[0], (None, None), m, sh(90), v[1], 250, w, 10,
                                                 'concat', init.Orthogonal(), True)
        return net, v, m, w, [1144, 90], True, 10
    if name == 'adenet_v3':                     # config 3 (README-era trimodal), modelzoo/adenet_v3.py:64: H = 500 LSTMs
        v = [T.tensor3('raw'), T.tensor3('dct'), T.tensor3('diff')]
        net, _ = modelzoo.adenet_v3.create_model(_DBN(make_ae(rng, 1200)), _DBN(make_ae(rng, 1200)), sh(1200), v[0],
                                                 (None, None), m, sh(90), v[1], sh(1200), v[2], 250, w, 26, 'sum')
        return net, v, m, w, [1200, 90, 1200], False, 26
    raise KeyError(name)


def other_configs(steps, mode):
    """BASELINE configs 1-3 at the reference's own batch and at 512, device-resident inputs, one B200."""
    import torch
    from ipavsr_b200 import layers as L
    from ipavsr_b200.function import function, tensor as T
    from ipavsr_b200.custom.objectives import temporal_softmax_loss, categorical_crossentropy
    from ipavsr_b200.custom import updates as U
    res = []
    cases = [('deltanet', 'config 1: deltanet D=1200 BLSTM-250 C=26, mean categorical cross-entropy, Adam', (26, 512), 'adam'),
             ('adenet_v1', 'config 2: adenet_v1 D=1144 + DCT 90, BatchNorm, BLSTM-250 -> BLSTM-500, C=10, Adam', (10, 512), 'adam'),
             ('adenet_v2', 'config 2: adenet_v2 concat D=1144 + DCT 90, LSTM-250 x2 -> BLSTM-250, C=10, Adam', (10, 512), 'adam'),
             ('adenet_v3', 'config 3: adenet_v3 sum, raw+DCT+diff, dropout, LSTM-500 x3 -> BLSTM-500 (FFMA recurrence: '
                           'H > 256), C=26, Adadelta lr 2.0', (26, 512), 'adadelta')]
    for name, desc, batches, rule in cases:
        net, v, m, w, dims, frame_level, classes = build_config_net(name, 1234)
        pred = L.get_output(net, deterministic=False)
        params = L.get_all_params(net, trainable=True)
        if frame_level:
            tg = T.imatrix('t')
            cost = temporal_softmax_loss(pred, tg, m)
        else:
            tg = T.ivector('t')
            cost = T.mean(categorical_crossentropy(pred, tg))
        upd = U.adam(cost, params, learning_rate=1e-3) if rule == 'adam' else U.adadelta(cost, params, learning_rate=2.0)
        train = function([v[0], tg, m] + v[1:] + [w], cost, updates=upd, gemm_mode=mode)
        for nb in batches:
            xs, mask, y = synth_batch(nb, 1235, dims=dims, classes=classes, frame_targets=frame_level)
            dx = [torch.from_numpy(x).cuda() for x in xs]
            dm, dy = torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()
            fn = lambda: train(dx[0], dy, dm, *dx[1:], THETA)
            for _ in range(6):        # (the early loss read-back lets the host run a step ahead: the caching allocator takes
                fn()                  #  a few more calls to reach the block set it then cycles through)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res.append({'config': desc, 'builder': name, 'utterances': nb, 'ms_per_step': ms, 'utterances_per_s': nb / ms * 1e3,
                        'loss_finite': bool(np.isfinite(loss))})
            del dx
        del train, net
        torch.cuda.empty_cache()
    return res


def inference_4stream(total, world, rank, steps, mode, timed):
    """BASELINE config 5: adenet_4stream (4 x 1200-px streams, concat, peepholes, H=250, C=26), `total` utterances x T=40,
    lengths U{10..40}, sharded total/world per GPU, no collective; inputs resident in HBM, probabilities read back."""
    import torch
    from ipavsr_b200 import modelzoo, init, layers as L
    from ipavsr_b200.function import function, tensor as T
    rng = np.random.default_rng(1236)
    np.random.seed(1236)
    aes = [make_ae(rng, 1200) for _ in range(4)]
    v = [T.tensor3('s%d' % (i + 1)) for i in range(4)]
    m, w = T.matrix('mask', dtype='uint8'), T.iscalar('theta')
    sh = (None, None, 1200)
    net, _ = modelzoo.adenet_4stream.create_model(aes[0], aes[1], aes[2], aes[3], sh, v[0], sh, v[1], sh, v[2], sh, v[3],
                                                  (None, None), m, H_LSTM, w, N_CLASSES, 'concat', init.Orthogonal(), True)
    val_fn = function([v[0], v[1], v[2], v[3], m, w], L.get_output(net, deterministic=True), gemm_mode=mode)
    n_local = total // world
    chunk = min(1024, n_local)
    lens = np.random.default_rng(77 + rank).integers(10, T_FRAMES + 1, size=n_local)
    mask = (np.arange(T_FRAMES)[None, :] < lens[:, None]).astype('uint8')
    dmask = torch.from_numpy(mask).cuda()
    xs = [torch.randn(n_local, T_FRAMES, 1200, device='cuda') * dmask[:, :, None] for _ in range(4)]
    masks = [dmask[c0:c0 + chunk].contiguous() for c0 in range(0, n_local, chunk)]

    def run():
        for ci, c0 in enumerate(range(0, n_local, chunk)):
            val_fn(xs[0][c0:c0 + chunk], xs[1][c0:c0 + chunk], xs[2][c0:c0 + chunk], xs[3][c0:c0 + chunk], masks[ci], THETA)
    ms, _ = timed(run, steps, 2)
    frames_valid = int(lens.sum())
    return {'workload': 'adenet_4stream inference (BASELINE config 5): %d utterances x T=40 x 4 streams of 1200 px, lengths '
                        'U{10..40}, %d per GPU in chunks of %d, no collective' % (total, n_local, chunk),
            'ms_per_pass': ms, 'utterances_per_s': total / ms * 1e3, 'frames_per_s': total * T_FRAMES / ms * 1e3,
            'valid_frames_per_s_per_gpu': frames_valid / ms * 1e3,
            'encoder_tflops_reference_arithmetic': total * T_FRAMES * 4 * 2.0 * (1200 * 2000 + 2000 * 1000 + 1000 * 500 + 500 * 50)
            / ms / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--batch', type=int, default=512, help='utterances per GPU per step (SURVEY 8d quotes config 3 at 512)')
    ap.add_argument('--mode', default=os.environ.get('IPAVSR_GEMM_MODE', 'f16x3'),
                    help='GEMM arithmetic: f16x3 (fp32-parity 3-product fp16 tensor cores, engine default) | tf32x3 | fp32 '
                         '(CUDA cores) | tf32 (single pass)')
    ap.add_argument('--cpu-sample', type=int, default=26)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip rooflines[], configs[] and inference_4stream')
    ap.add_argument('--no-prefetch', action='store_true', help='e2e without the double-buffered input prefetch')
    ap.add_argument('--prefetch-depth', type=int, default=int(os.environ.get('IPAVSR_BENCH_PREFETCH_DEPTH', '1')),
                    help='e2e: how many steps ahead the upload of a batch is staged (1 = the next step)')
    ap.add_argument('--only-e2e', action='store_true', help='experiment aid: print the device-resident and e2e step times only')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    metric = 'AdeNet-v2 trimodal train utterances/s'
    mb = args.batch * T_FRAMES * sum(STREAM_DIMS) * 4 // 1000000
    config = {'workload': 'adenet_3stream train step (raw1200+diff1200+dct90 -> DBNF 2000-1000-500-50 -> delta(9) -> '
                          'LSTM-250 x3 -> concat -> BLSTM-250 -> softmax-26), T=40, lengths U{12..40}, Adam',
              'utterances_per_gpu': args.batch, 'global_batch': args.batch * world, 'frames': T_FRAMES,
              'parallelism': 'dp%d' % world,
              'timing': '4 distinct batches in rotation; inputs (%d MB/step/GPU padded) larger than the 126 MB L2' % mb,
              'cpu_arm_batch': min(args.cpu_sample, args.batch)}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        import torch
        # all the host threads the BLAS behind NumPy can use (torchrun exports OMP_NUM_THREADS=1 for its workers)
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=ncpu)
        except Exception:
            pass
        torch.set_num_threads(ncpu)
        n = min(args.cpu_sample, args.batch)
        steps = max(1, min(args.steps, 3))
        rate, spstep = cpu_reference_step_rate(n, steps, 1)
        line = {'impl': 'reference', 'metric': metric, 'value': rate, 'unit': 'utterances/s', 'n_gpus': args.gpus,
                'steps': steps, 'warmup': 1, 'ms_per_step': spstep * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': rate, 'unit': 'utterances/s', 'cores': ncpu, 'kind': 'port',
                                 'sample': '%d-utterance steps of the same workload (NumPy float32 oracle, BLAS dots); '
                                           'Theano/Lasagne are not installable here' % n},
                'e2e': {'value': rate, 'unit': 'utterances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return 0

    import torch
    from ipavsr_b200 import layers as L, _lib
    from ipavsr_b200.engine import get_engine
    from ipavsr_b200.function import function, tensor as T
    from ipavsr_b200.custom.objectives import temporal_softmax_loss
    from ipavsr_b200.custom.updates import adam
    from ipavsr_b200.derived import DiffImages, DctFeatures

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    net, v, mask_var, window = build_network()
    eng = get_engine(net, gemm_mode=args.mode)
    if world > 1:
        from ipavsr_b200 import parallel
        parallel.attach(eng)
    targets = T.imatrix('targets')
    pred = L.get_output(net, deterministic=False)
    cost = temporal_softmax_loss(pred, targets, mask_var)
    params = L.get_all_params(net, trainable=True)
    train = function([v[0], v[1], v[2], targets, mask_var, window], cost, updates=adam(cost, params, learning_rate=1e-3))
    lib = _lib.load()

    NB = 4                                       # distinct batches in rotation (different data and lengths)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    host, dev = [], []
    for b in range(NB):
        xs, mask, y = synth_batch(args.batch, 1000 + 17 * rank + b)
        # the host side keeps what a runner holds: the raw stream, the mask and the targets; the device side the three
        # padded streams of the reference's call convention
        host.append((pin(xs[0]), pin(mask), pin(y), int(mask.sum())))
        dev.append(([torch.from_numpy(x).cuda() for x in xs], torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()))
        del xs
    valid_rows = [h[3] for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = lib.ipavsr_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, lib.ipavsr_launch_count() - l0

    it = [0]

    def step_dev():
        xs, m, y = dev[it[0] % NB]
        it[0] += 1
        return train(xs[0], xs[1], xs[2], y, m, THETA)

    with ClockSampler(local) as clk:
        ms_dev, launches = timed(step_dev, args.steps, max(args.warmup, 3))

    # ---- e2e: pinned host raw stream in, loss out, every step; diff / DCT streams derived on the device ----
    def e2e_args(i):
        raw, m, y, _ = host[i % NB]
        return (raw, DiffImages(raw), DctFeatures(raw, IMAGE_SHAPE, DCT_COEFF), y, m, THETA)
    cache = [e2e_args(i) for i in range(NB)]
    jt = [0]
    depth = max(1, min(args.prefetch_depth, NB - 1))
    if not args.no_prefetch:
        for k in range(depth):
            train.prefetch(*cache[k % NB])

    def step_e2e():
        i = jt[0]
        jt[0] += 1
        if not args.no_prefetch:
            # the upload of a later step is staged by this call once its own kernels are enqueued (before it reads the loss
            # back): the upload overlaps this step's compute and the staging work hides behind it
            train.prefetch(*cache[(i + depth) % NB], defer=True)
        return train(*cache[i % NB])
    ms_e2e, _ = timed(step_e2e, args.steps, 3)
    eng._prefetched = []
    if args.only_e2e:
        if rank == 0:
            print(json.dumps({'n_gpus': world, 'prefetch_depth': depth, 'ms_dev': ms_dev, 'ms_e2e': ms_e2e,
                              'value': args.batch * world / (ms_dev * 1e-3), 'e2e': args.batch * world / (ms_e2e * 1e-3)}))
        if world > 1:
            dist.destroy_process_group()
        return 0
    # bytes copied host -> device per step: the valid frames of the raw stream, the targets, and the plan's index tables
    # (pack M+1, valid M, unpack / perm / unperm N*T each, order / inv N each, int32; sorted mask N*T bytes; offsets N+1
    # int64) — counted every step although a plan is re-used while the same lengths come back
    mv, nt = int(np.mean(valid_rows)), args.batch * T_FRAMES
    plan_bytes = 4 * (2 * mv + 1 + 3 * nt + 2 * args.batch) + nt + 8 * (args.batch + 1)
    h2d = mv * STREAM_DIMS[0] * 4 + nt * 4 + plan_bytes
    value = args.batch * world / (ms_dev * 1e-3)
    e2e = args.batch * world / (ms_e2e * 1e-3)

    # ---- e2e in the reference's own call convention: three padded host streams uploaded every step ----
    hp = []
    for b in range(2):
        xs, mask, y = synth_batch(args.batch, 3000 + 17 * rank + b)
        hp.append(([pin(x) for x in xs], pin(mask), pin(y)))
        del xs
    kt = [0]
    train.prefetch(hp[0][0][0], hp[0][0][1], hp[0][0][2], hp[0][2], hp[0][1], THETA)

    def step_e2e_padded():
        i = kt[0]
        kt[0] += 1
        nx = hp[(i + 1) % 2]
        train.prefetch(nx[0][0], nx[0][1], nx[0][2], nx[2], nx[1], THETA, defer=True)
        cx = hp[i % 2]
        return train(cx[0][0], cx[0][1], cx[0][2], cx[2], cx[1], THETA)
    ms_e2e_p, _ = timed(step_e2e_padded, max(4, args.steps // 2), 2)
    eng._prefetched = []
    del hp

    # ---- the same step at 960 utterances per GPU (round-1 headline batch), device-resident ----
    ms_alt = None
    alt = 960 if args.batch != 960 else 512
    xs5, mask5, y5 = synth_batch(alt, 2000 + rank)
    d5 = [torch.from_numpy(x).cuda() for x in xs5]
    m5, y5d = torch.from_numpy(mask5).cuda(), torch.from_numpy(y5).cuda()
    ms_alt, _ = timed(lambda: train(d5[0], d5[1], d5[2], y5d, m5, THETA), max(4, args.steps // 2), 3)
    del d5, xs5
    # ---- the reference's own batch (26 utterances, avletters/trimodal.py:356-359): launch-bound regime ----
    xs26, mask26, y26 = synth_batch(26, 2600 + rank)
    d26 = [torch.from_numpy(x).cuda() for x in xs26]
    m26, y26d = torch.from_numpy(mask26).cuda(), torch.from_numpy(y26).cuda()
    ms_26, launches_26 = timed(lambda: train(d26[0], d26[1], d26[2], y26d, m26, THETA), args.steps, 3)
    # ---- forward-only (deterministic) pass of the same network: frames/s ----
    val_fn = function([v[0], v[1], v[2], mask_var, window], L.get_output(net, deterministic=True))
    ft = [0]

    def step_fwd():
        xs, m, y = dev[ft[0] % NB]
        ft[0] += 1
        return val_fn(xs[0], xs[1], xs[2], m, THETA)
    ms_fwd, _ = timed(step_fwd, max(3, args.steps // 2), 3)
    fwd_frames = args.batch * world * T_FRAMES / (ms_fwd * 1e-3)
    del dev

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    # ---- roofline of the dominant kernel: the fc1 encoder GEMM on the packed rows (valid frames + the zero row) ----
    roofline = None
    cpu_baseline = None
    rooflines = None
    configs = None
    if rank == 0:
        M, K, N = int(np.mean(valid_rows)) + 1, STREAM_DIMS[0], ENC[0]
        A = torch.randn(M, K, device='cuda')
        B = torch.randn(K, N, device='cuda')
        Cm = torch.empty(M, N, device='cuda')
        bias = torch.zeros(N, device='cuda')
        mode = {'fp32': 0, 'tf32x3': 1, 'tf32': 2, 'f16x3': 4}[args.mode]
        st = _stream()
        flush = torch.empty(192 * 1024 * 1024 // 4, device='cuda')
        if mode == 1:
            ah, al, bh, bl = torch.empty_like(A), torch.empty_like(A), torch.empty_like(B), torch.empty_like(B)
            _lib.call('ipavsr_tf32_split_rna', A.data_ptr(), ah.data_ptr(), al.data_ptr(), A.numel(), st)
            _lib.call('ipavsr_tf32_split_rna', B.data_ptr(), bh.data_ptr(), bl.data_ptr(), B.numel(), st)
            run = lambda: _lib.call('ipavsr_gemm_tf32x3_presplit', 0, 0, M, N, K, ah.data_ptr(), al.data_ptr(), K,
                                    bh.data_ptr(), bl.data_ptr(), N, Cm.data_ptr(), N, bias.data_ptr(), 1, 0, None, None, st)
        elif mode == 4:
            # inside the step the operands arrive as fp16 hi/lo + per-tensor scale exponents
            ah, al = torch.empty(M, K, dtype=torch.float16, device='cuda'), torch.empty(M, K, dtype=torch.float16, device='cuda')
            bh, bl = torch.empty(K, N, dtype=torch.float16, device='cuda'), torch.empty(K, N, dtype=torch.float16, device='cuda')
            sc = torch.zeros(4, device='cuda')
            _lib.call('ipavsr_f16_split', A.data_ptr(), K, M, K, ah.data_ptr(), al.data_ptr(), K, sc.data_ptr(),
                      sc.data_ptr() + 4, 0, st)
            _lib.call('ipavsr_f16_split', B.data_ptr(), N, K, N, bh.data_ptr(), bl.data_ptr(), N, sc.data_ptr() + 8,
                      sc.data_ptr() + 12, 0, st)
            run = lambda: _lib.call('ipavsr_gemm_f16x3', 0, 0, M, N, K, ah.data_ptr(), al.data_ptr(), K, sc.data_ptr() + 4,
                                    bh.data_ptr(), bl.data_ptr(), N, sc.data_ptr() + 12, Cm.data_ptr(), N, bias.data_ptr(),
                                    1, 0, None, None, None, 0, st)
        else:
            run = lambda: _lib.call('ipavsr_gemm', mode, 0, 0, M, N, K, A.data_ptr(), K, B.data_ptr(), N, Cm.data_ptr(), N,
                                    bias.data_ptr(), 1, 0, None, 0, st)
        gemm_ms = time_kernel(run, flush, reps=10, warm=3)
        peak = float(peaks.get('bf16_tflops', 1590.0))
        achieved = 2.0 * M * N * K / (gemm_ms * 1e-3) / 1e12
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (profiles/r02c_ncu_gemm_summary.md):
        # same N and K, and M within 5 % of the rows timed here (the packed row count follows the random lengths)
        traffic, traffic_at, ncu_ent = None, None, None
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'r02_dominant_kernel.json')))
            for ent in prof.get('entries', [prof]):
                sh = ent.get('shape') or [0, 0, 0]
                if ent.get('mode') == args.mode and sh[1:] == [N, K] and abs(sh[0] - M) <= 0.05 * M:
                    traffic, traffic_at = ent.get('dram_bytes_per_launch'), sh
                    ncu_ent = ent
        except Exception:
            pass
        p_before = _lib.load().ipavsr_debug_gemm_persistent_launches()
        run()
        kname = 'gemm_f16p_kernel (persistent)' if _lib.load().ipavsr_debug_gemm_persistent_launches() > p_before else 'gemm_tc_kernel'
        roofline = {'bound': 'tensor', 'kernel': '%s: encoder fc1 GEMM %dx%dx%d (%s) on the packed rows of a '
                                                  '%d-utterance batch' % (kname, M, N, K, args.mode, args.batch),
                    'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                    'traffic_measured_at_shape': traffic_at,
                    'peak_source': 'MEASURED_PEAKS.json bf16 burst' if peaks else 'fallback 1.59 PFLOP/s',
                    'note': {'f16x3': 'algorithmic FLOPs; fp32 parity on 16-bit tensor cores costs 3 MMAs per product, so the '
                                      'ceiling of this mode is 1/3 of the bf16 peak',
                             'tf32x3': 'algorithmic FLOPs; fp32 parity on tf32 tensor cores costs 3 MMAs/product and tf32 peak '
                                       'is half of bf16, so the ceiling of this mode is 1/6 of the bf16 peak'}.get(
                                 args.mode, 'algorithmic FLOPs'),
                    'ms_per_launch': gemm_ms}
        if ncu_ent is not None:
            roofline['ncu'] = {'kernel': ncu_ent.get('kernel'), 'tensor_pipe_active_pct': ncu_ent.get('tensor_pipe_active_pct'),
                               'sm_clock_ghz_during_kernel': ncu_ent.get('sm_clock_ghz'), 'time_us': ncu_ent.get('time_us'),
                               'source': 'profiles/r02c_ncu_gemm_summary.md (committed ncu --set full capture, not this run)'}
        del A, B, Cm, flush
        if not args.no_extras and world == 1:
            hbm = float(peaks.get('hbm_gbs', 6650.0))
            rooflines = streaming_rooflines(hbm, 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s')
            torch.cuda.empty_cache()
            configs = other_configs(max(5, args.steps // 2), args.mode)
            torch.cuda.empty_cache()
    inference = None
    if not args.no_extras:
        inference = inference_4stream(4096, world, rank, 3, args.mode, timed)
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        rate, spstep = cpu_reference_step_rate(args.cpu_sample, 2, 1)
        cpu_baseline = {'value': rate, 'unit': 'utterances/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                        'sample': '%d-utterance training steps of the same workload, NumPy float32 oracle '
                                  '(%.1f s/step)' % (args.cpu_sample, spstep)}
    if rank == 0:
        line = {'metric': metric, 'value': value, 'unit': 'utterances/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32' if args.mode == 'fp32' else args.mode, 'data': 'synthetic',
                'config': config, 'clocks': clk.summary(),
                'e2e': {'value': e2e, 'unit': 'utterances/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 12,
                        'ms_per_step': ms_e2e,
                        'inputs': 'pinned host raw stream (valid frames only, one batched copy-engine transfer per step: '
                                  'ipavsr_upload_ragged), targets and index tables; diff-image and DCT+delta streams '
                                  'computed on the device from it (ipavsr_b200.derived); loss read back'},
                'e2e_padded_streams': {'value': args.batch * world / (ms_e2e_p * 1e-3), 'unit': 'utterances/s',
                                       'ms_per_step': ms_e2e_p, 'h2d_bytes_per_step': args.batch * T_FRAMES * (sum(STREAM_DIMS) * 4 + 4),
                                       'note': "the reference's call convention: three padded host streams per step"},
                'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu_baseline,
                'model_tflops': flops_per_utt_train() * args.batch * world / (ms_dev * 1e-3) / 1e12,
                'fwd_frames_per_s': fwd_frames, 'fwd_ms_per_batch': ms_fwd,
                'batch_%d' % alt: {'value': alt * world / (ms_alt * 1e-3), 'unit': 'utterances/s', 'ms_per_step': ms_alt,
                                   'note': 'device-resident step at %d utterances per GPU' % alt},
                'batch_26': {'value': 26 * world / (ms_26 * 1e-3), 'unit': 'utterances/s', 'ms_per_step': ms_26,
                             'gpu_launches_per_step': int(launches_26) // max(args.steps, 1),
                             'note': "the reference's own batch (avletters/trimodal.py:356-359); launch-bound when run "
                                     "eagerly, so forward + loss + backward replay one CUDA graph per step"},
                'rooflines': rooflines, 'inference_4stream': inference, 'configs': configs}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
