"""BASELINE.json config 5: adenet_4stream large-batch inference (4 x 1200-pixel streams, DBNF 2000-1000-500-50, delta(9),
LSTM-250 with peepholes, concat fusion, BLSTM-250, per-frame softmax-26), T=40, variable lengths, sharded by utterance with
no collective.      python tools/bench_inference.py [--batch 4096] [--chunk 1024] [--steps 5]
Inputs are resident in HBM; utterances are processed in chunks (activations of 4 encoders at 4096 x 40 frames would not be
needed at once).  Prints utterances/s and frames/s for this GPU."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ipavsr_b200 import modelzoo, nonlinearities as nl, init, layers as L
from ipavsr_b200.function import function, tensor as T

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=4096)
ap.add_argument('--chunk', type=int, default=1024)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--mode', default='f16x3')
args = ap.parse_args()
os.environ['IPAVSR_GEMM_MODE'] = args.mode
rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
rng = np.random.default_rng(1236)
np.random.seed(1236)
ENC, Tn, H, C = (2000, 1000, 500, 50), 40, 250, 26
aes = []
for _ in range(4):
    s = (1200,) + ENC
    W = [rng.normal(0, 1.0 / np.sqrt(s[i]), (s[i], s[i + 1])).astype('float32') for i in range(4)]
    b = [rng.normal(0, 0.1, (s[i + 1],)).astype('float32') for i in range(4)]
    aes.append((W, b, list(ENC), [nl.select_nonlinearity(a) for a in ('sigmoid', 'sigmoid', 'sigmoid', 'linear')]))
v = [T.tensor3('s%d' % (i + 1)) for i in range(4)]
mask_var, window = T.matrix('mask', dtype='uint8'), T.iscalar('theta')
sh = (None, None, 1200)
net, _ = modelzoo.adenet_4stream.create_model(aes[0], aes[1], aes[2], aes[3], sh, v[0], sh, v[1], sh, v[2], sh, v[3],
                                              (None, None), mask_var, H, window, C, 'concat', init.Orthogonal(), True)
val_fn = function([v[0], v[1], v[2], v[3], mask_var, window], L.get_output(net, deterministic=True))
n_local = args.batch // world
chunk = min(args.chunk, n_local)
lens = rng.integers(10, Tn + 1, size=chunk)
mask = (np.arange(Tn)[None, :] < lens[:, None]).astype('uint8')
dmask = torch.from_numpy(mask).cuda()
xs = [torch.randn(chunk, Tn, 1200, device='cuda') * dmask[:, :, None] for _ in range(4)]


def run():
    for _ in range((n_local + chunk - 1) // chunk):
        val_fn(xs[0], xs[1], xs[2], xs[3], dmask, 9)       # probabilities are read back to the host per chunk


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(json.dumps({'workload': 'adenet_4stream inference, %d utterances/GPU in chunks of %d, T=40' % (n_local, chunk),
                  'mode': args.mode, 'n_gpus': world, 'ms_per_pass': ms, 'utterances_per_s_per_gpu': n_local / ms * 1e3,
                  'frames_per_s_per_gpu': n_local * Tn / ms * 1e3}))
