#!/usr/bin/env python
"""Benchmark of the AdeNet hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch B] [--mode f16x3|tf32x3|fp32|tf32]

Workload (config.workload): the AdeNet-v2 late-fusion *trimodal* network (`modelzoo.adenet_3stream`:
raw 1200-px ROI + diff-image 1200-px + DCT 90, each through a DBNF encoder 2000-1000-500-50 -> DeltaLayer(theta=9) ->
masked LSTM-250 with peepholes, concat fusion, BLSTM-250 aggregate, per-frame softmax over 26 classes), one
training step = forward + temporal_softmax_loss + full backward + Adam, T=40 padded frames, variable lengths,
synthetic data, random-init weights.  Per-GPU batch is fixed (weak scaling); with N>1 one process per GPU
(torchrun) shards utterances and all-reduces the flat gradient arena over NCCL.

One JSON line on stdout (rank 0).  `value` = utterances/s with inputs already resident in HBM; `e2e` = the same
step through the public `function(...)` callable with pinned HOST inputs (H2D inside the timed region) and the
loss read back (D2H) every step.  `--impl reference` times the CPU restatement of the reference path (the NumPy
oracle — Theano/Lasagne cannot be installed here) on the host cores instead.
"""
import argparse
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
ENC = (2000, 1000, 500, 50)
ENC_ACTS = ('sigmoid', 'sigmoid', 'sigmoid', 'linear')


def build_network(seed=1234):
    from ipavsr_b200 import modelzoo, nonlinearities as nl, init
    from ipavsr_b200.function import tensor as T
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    aes = []
    for D in STREAM_DIMS:
        s = (D,) + ENC
        W = [rng.normal(0, 1.0 / np.sqrt(s[i]), (s[i], s[i + 1])).astype('float32') for i in range(4)]
        b = [rng.normal(0, 0.1, (s[i + 1],)).astype('float32') for i in range(4)]
        aes.append((W, b, list(ENC), [nl.select_nonlinearity(a) for a in ENC_ACTS]))
    v = [T.tensor3('s%d' % (i + 1)) for i in range(3)]
    mask = T.matrix('mask', dtype='uint8')
    window = T.iscalar('theta')
    net, fuse = modelzoo.adenet_3stream.create_model(
        aes[0], aes[1], aes[2], (None, None, STREAM_DIMS[0]), v[0], (None, None, STREAM_DIMS[1]), v[1],
        (None, None, STREAM_DIMS[2]), v[2], (None, None), mask, H_LSTM, window, N_CLASSES, 'concat',
        init.Orthogonal(), True)
    return net, v, mask, window


def synth_batch(n, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(12, T_FRAMES + 1, size=n)
    lens[0] = T_FRAMES
    mask = (np.arange(T_FRAMES)[None, :] < lens[:, None]).astype('uint8')
    xs = []
    for D in STREAM_DIMS:
        x = rng.standard_normal((n, T_FRAMES, D), dtype=np.float32)
        x *= mask[:, :, None]
        xs.append(x)
    y = np.repeat(rng.integers(0, N_CLASSES, size=(n, 1)), T_FRAMES, 1).astype('int32')
    return xs, mask, y


def flops_per_utt_train():
    """Algorithmic FLOPs of one training step per utterance (fwd + dgrad + wgrad; first-layer dgrad not needed)."""
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
            self.stop.wait(0.2)

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--batch', type=int, default=960,
                    help='utterances per GPU per step (default 960 = 2 x 15 x 32: the tensor-core LSTM works on 32-utterance '
                         'tiles, 15 of its 8-CTA clusters are co-resident, so multiples of 480 fill whole waves)')
    ap.add_argument('--mode', default=os.environ.get('IPAVSR_GEMM_MODE', 'f16x3'),
                    help='GEMM arithmetic: f16x3 (fp32-parity 3-product fp16 tensor cores, default) | tf32x3 (fp32-parity '
                         '3xTF32) | fp32 (CUDA cores) | tf32 (single pass)')
    ap.add_argument('--cpu-sample', type=int, default=26)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-prefetch', action='store_true', help='e2e without the double-buffered input prefetch')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    metric = 'AdeNet-v2 trimodal train utterances/s'
    config = {'workload': 'adenet_3stream train step (raw1200+diff1200+dct90 -> DBNF 2000-1000-500-50 -> delta(9) -> '
                          'LSTM-250 x3 -> concat -> BLSTM-250 -> softmax-26), T=40, variable lengths, Adam',
              'utterances_per_gpu': args.batch, 'global_batch': args.batch * world, 'frames': T_FRAMES,
              'parallelism': 'dp%d' % world, 'timing': 'inputs (%d MB/step/GPU) larger than the 126 MB L2' % (args.batch * T_FRAMES * sum(STREAM_DIMS) * 4 // 1000000)}

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
        cores = ncpu
        n = min(args.cpu_sample, args.batch)
        steps = max(1, min(args.steps, 3))
        rate, spstep = cpu_reference_step_rate(n, steps, 1)
        line = {'impl': 'reference', 'metric': metric, 'value': rate, 'unit': 'utterances/s', 'n_gpus': args.gpus,
                'steps': steps, 'warmup': 1, 'ms_per_step': spstep * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': rate, 'unit': 'utterances/s', 'cores': cores, 'kind': 'port',
                                 'sample': '%d-utterance steps of the same workload (NumPy float32 oracle, BLAS dots); '
                                           'Theano/Lasagne are not installable here' % n},
                'e2e': {'value': rate, 'unit': 'utterances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return 0

    import torch
    import ctypes as C
    from ipavsr_b200 import layers as L, _lib
    from ipavsr_b200.engine import get_engine
    from ipavsr_b200.function import function, tensor as T
    from ipavsr_b200.custom.objectives import temporal_softmax_loss
    from ipavsr_b200.custom.updates import adam

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    os.environ['IPAVSR_GEMM_MODE'] = args.mode
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

    xs, mask, y = synth_batch(args.batch, 1000 + rank)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    hx = [pin(x) for x in xs]
    hmask, hy = pin(mask), pin(y)
    dx = [h.cuda(non_blocking=True) for h in hx]
    dmask, dy = hmask.cuda(), hy.cuda()
    lib = _lib.load()

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

    step_dev = lambda: train(dx[0], dx[1], dx[2], dy, dmask, THETA)
    step_e2e = lambda: train(hx[0], hx[1], hx[2], hy, hmask, THETA)
    with ClockSampler(local) as clk:
        ms_dev, launches = timed(step_dev, args.steps, max(args.warmup, 3))
    # e2e: every step copies its inputs from pinned host memory and reads the loss back.  The copy of step i+1 is started
    # (train.prefetch, the public double-buffering call) before step i is launched, so that it overlaps step i's compute;
    # the first prefetch happens before the warm-up.
    e2e_args = (hx[0], hx[1], hx[2], hy, hmask, THETA)
    train.prefetch(*e2e_args)

    def step_e2e_pf():
        train.prefetch(*e2e_args)
        return train(*e2e_args)
    ms_e2e, _ = timed(step_e2e_pf if not args.no_prefetch else step_e2e, args.steps, 3)
    value = args.batch * world / (ms_dev * 1e-3)
    e2e = args.batch * world / (ms_e2e * 1e-3)
    # the same step fed by device-side batch assembly: the packed variable-length dataset stays in HBM
    # and every step gathers its padded (N, T, F) streams + mask with ipavsr_batch_gather (utils/datagen.DeviceDataset)
    from ipavsr_b200.utils import datagen as DG
    lens = mask.sum(axis=1).astype(np.int64)
    packed = [np.concatenate([x[i, :lens[i]] for i in range(len(lens))] * 2, axis=0) for x in xs]
    dsets = [DG.DeviceDataset(p, np.concatenate([lens, lens])) for p in packed]
    del packed
    rng_idx = np.random.default_rng(7 + rank)
    idx_lists = [rng_idx.permutation(2 * len(lens))[:args.batch] for _ in range(8)]
    ds_step = [0]

    def step_dataset():
        idx = idx_lists[ds_step[0] % len(idx_lists)]
        ds_step[0] += 1
        x0, m0 = dsets[0].gather(idx, T_FRAMES, with_mask=True)
        return train(x0, dsets[1].gather(idx, T_FRAMES), dsets[2].gather(idx, T_FRAMES), dy, m0, THETA)
    ms_ds, _ = timed(step_dataset, args.steps, 3)
    del dsets
    # SURVEY 8(d) quotes config 3 at 512 utterances per GPU: the same device-resident step at that batch, for comparison
    # (16 LSTM tiles per launch = one full wave of the 15 co-resident clusters plus an almost empty one)
    ms_512 = None
    if args.batch != 512:
        xs5, mask5, y5 = synth_batch(512, 2000 + rank)
        d5 = [torch.from_numpy(x).cuda() for x in xs5]
        m5, y5d = torch.from_numpy(mask5).cuda(), torch.from_numpy(y5).cuda()
        ms_512, _ = timed(lambda: train(d5[0], d5[1], d5[2], y5d, m5, THETA), args.steps, 3)
        del d5, xs5
    # forward-only (deterministic) pass of the same network: frames/s
    val_fn = function([v[0], v[1], v[2], mask_var, window], L.get_output(net, deterministic=True))
    ms_fwd, _ = timed(lambda: val_fn(dx[0], dx[1], dx[2], dmask, THETA), max(3, args.steps // 2), 3)
    fwd_frames = args.batch * world * T_FRAMES / (ms_fwd * 1e-3)
    h2d = sum(int(h.numel() * h.element_size()) for h in hx) + int(hmask.numel()) + int(hy.numel() * 4)

    # ---- roofline of the dominant kernel: the fc1 encoder GEMM (M = batch*T rows, K=1200, N=2000) ----
    roofline = None
    cpu_baseline = None
    if rank == 0:
        M, K, N = args.batch * T_FRAMES, STREAM_DIMS[0], ENC[0]
        A = torch.randn(M, K, device='cuda')
        B = torch.randn(K, N, device='cuda')
        Cm = torch.empty(M, N, device='cuda')
        bias = torch.zeros(N, device='cuda')
        mode = {'fp32': 0, 'tf32x3': 1, 'tf32': 2, 'f16x3': 4}[args.mode]
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        flush = torch.empty(192 * 1024 * 1024 // 4, device='cuda')
        if mode == 1:
            # inside the step the operands arrive already split (weights once per step, activations by the producing
            # epilogue), so the dominant kernel is the presplit 3xTF32 tcgen05 GEMM
            ah, al, bh, bl = torch.empty_like(A), torch.empty_like(A), torch.empty_like(B), torch.empty_like(B)
            _lib.call('ipavsr_tf32_split_rna', A.data_ptr(), ah.data_ptr(), al.data_ptr(), A.numel(), st)
            _lib.call('ipavsr_tf32_split_rna', B.data_ptr(), bh.data_ptr(), bl.data_ptr(), B.numel(), st)
            run = lambda: _lib.call('ipavsr_gemm_tf32x3_presplit', 0, 0, M, N, K, ah.data_ptr(), al.data_ptr(), K,
                                    bh.data_ptr(), bl.data_ptr(), N, Cm.data_ptr(), N, bias.data_ptr(), 1, 0, None, None, st)
        elif mode == 4:
            # likewise for the fp16 three-product mode: operands arrive as fp16 hi/lo + per-tensor scale exponents
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
        for _ in range(3):
            run()
        tot = 0.0
        reps = 10
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        gemm_ms = tot / reps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('bf16_tflops', 1590.0))
        achieved = 2.0 * M * N * K / (gemm_ms * 1e-3) / 1e12
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'r01_dominant_kernel.json')))
            for ent in prof.get('entries', [prof]):
                if ent.get('mode') == args.mode and ent.get('shape') == [M, N, K]:
                    traffic = ent.get('dram_bytes_per_launch')
        except Exception:
            pass
        roofline = {'bound': 'tensor', 'kernel': 'gemm_tc_kernel: encoder fc1 GEMM %dx%dx%d (%s)' % (M, N, K, args.mode),
                    'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                    'peak_source': 'MEASURED_PEAKS.json bf16 burst' if peaks else 'fallback 1.59 PFLOP/s',
                    'note': {'f16x3': 'algorithmic FLOPs; fp32 parity on 16-bit tensor cores costs 3 MMAs per product, so the '
                                      'ceiling of this mode is 1/3 of the bf16 peak',
                             'tf32x3': 'algorithmic FLOPs; fp32 parity on tf32 tensor cores costs 3 MMAs/product and tf32 peak '
                                       'is half of bf16, so the ceiling of this mode is 1/6 of the bf16 peak'}.get(
                                 args.mode, 'algorithmic FLOPs'),
                    'ms_per_launch': gemm_ms}
        if not args.no_cpu_baseline and world == 1:
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
                        'ms_per_step': ms_e2e},
                'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu_baseline,
                'model_tflops': flops_per_utt_train() * args.batch * world / (ms_dev * 1e-3) / 1e12,
                'fwd_frames_per_s': fwd_frames, 'fwd_ms_per_batch': ms_fwd,
                'batch_512': (None if ms_512 is None else
                              {'value': 512 * world / (ms_512 * 1e-3), 'unit': 'utterances/s', 'ms_per_step': ms_512,
                               'note': 'device-resident step at 512 utterances per GPU (SURVEY 8d batch)'}),
                'device_dataset': {'value': args.batch * world / (ms_ds * 1e-3), 'unit': 'utterances/s',
                                   'ms_per_step': ms_ds,
                                   'note': 'batches gathered on the device from a packed dataset resident in HBM '
                                           '(ipavsr_batch_gather), no host copy in the step'}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
