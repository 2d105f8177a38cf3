"""Auto-encoder fine-tuning throughput: the reference's 8-layer DBNF auto-encoder
(1200-2000-1000-500-50-500-1000-2000-1200, sigmoid / linear bottleneck and output; `avletters/trimodal.py:41-89`) trained
through the nolearn-style NeuralNet mirror — squared error + 0.005 * L2, Nesterov momentum — on synthetic frames.
    python tools/bench_finetune.py [--batch 128,4096,32768] [--steps 20] [--json out.json]
Reports frames/s of the training step (host batch -> device -> loss read back, like NeuralNet.fit) and model TFLOP/s
(6 * parameters flop per frame minus the first layer's dgrad)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument('--batch', default='128,4096,32768')
ap.add_argument('--steps', type=int, default=20)
ap.add_argument('--json', default=None)
ap.add_argument('--profile', action='store_true', help='per-entry-point device time of one step (CUDA events per call)')
ap.add_argument('--mode', default='f16x3', help='GEMM arithmetic: f16x3 (tcgen05, fp32 parity) | tf32x3 | fp32 (FFMA)')
args = ap.parse_args()
os.environ['IPAVSR_GEMM_MODE'] = args.mode

from ipavsr_b200 import layers as L
from ipavsr_b200.nonlinearities import sigmoid, linear
from ipavsr_b200.custom.nolearn_net import NeuralNet
from ipavsr_b200.custom.objectives import squared_error
from ipavsr_b200.custom.updates import nesterov_momentum

sizes = (1200, 2000, 1000, 500, 50, 500, 1000, 2000, 1200)
rng = np.random.default_rng(0)
specs = [(L.InputLayer, {'name': 'input', 'shape': (None, sizes[0])})]
for i in range(8):
    specs.append((L.DenseLayer, {'name': 'output' if i == 7 else 'l%d' % (i + 1), 'num_units': sizes[i + 1],
                                 'nonlinearity': linear if i in (3, 7) else sigmoid,
                                 'W': (rng.normal(size=(sizes[i], sizes[i + 1])) / np.sqrt(sizes[i])).astype(np.float32),
                                 'b': np.zeros(sizes[i + 1], np.float32)}))
results = []
for B in [int(b) for b in args.batch.split(',')]:
    net = NeuralNet(layers=specs, max_epochs=1, objective_loss_function=squared_error, update=nesterov_momentum,
                    regression=True, update_learning_rate=0.001, update_momentum=0.05, objective_l2=0.005, batch_size=B)
    net.initialize()
    X = torch.randn(B, 1, sizes[0]).pin_memory().numpy()
    y = X.reshape(B, sizes[0])
    lib = net.train_iter_.engine.lib
    nparam = sum(sizes[i] * sizes[i + 1] for i in range(8))
    flop = B * (6.0 * nparam - 2.0 * sizes[0] * sizes[1])
    Xd = torch.from_numpy(X).cuda()
    yd = Xd.reshape(B, sizes[0])
    r = {'mode': args.mode, 'batch': B}
    # 'resident': what NeuralNet.fit does (the training matrix lives in HBM, batches are row views; the loss is read back
    # every step); 'host_batches': every step uploads its batch and targets from pinned host memory
    for tag, (a, b) in (('resident', (Xd, yd)), ('host_batches', (X, y))):
        for _ in range(3):
            loss = net.train_iter_(a, b)
        torch.cuda.synchronize()
        n0 = lib.ipavsr_launch_count()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss = net.train_iter_(a, b)           # returns the loss to the host: synchronises every step, like fit()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        r[tag] = {'ms_per_step': dt * 1e3, 'frames_per_s': B / dt, 'model_tflops': flop / dt / 1e12}
        r['launches_per_step'] = (lib.ipavsr_launch_count() - n0) / args.steps
    r['loss'] = float(loss)
    results.append(r)
    print(json.dumps(r), flush=True)
    if args.profile:
        from ipavsr_b200 import engine as E, _lib
        prof = E._Profiler(by_shape=True)
        orig = _lib.call
        _lib.call = E._lib.call = prof.call
        for _ in range(3):
            net.train_iter_(X, y)
        summ = prof.summary()
        _lib.call = E._lib.call = orig
        tot = sum(t for n, t in summ.values()) / 3
        print('  batch %d: sum of kernels %.3f ms/step (wall %.3f)' % (B, tot, dt * 1e3))
        for name, (n, t) in sorted(summ.items(), key=lambda kv: -kv[1][1])[:14]:
            print('    %-64s n/step=%4.1f %8.3f ms/step' % (name, n / 3, t / 3), flush=True)
    del net
if args.json:
    json.dump(results, open(args.json, 'w'), indent=1)
