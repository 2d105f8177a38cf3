"""Evaluation on the device: `ipavsr_vote_eval` and the `utils/evaluate.py` mirror against the oracle
(oracle/evaluate.py) and the reference's own results (tests/golden/evaluate.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import evaluate as OE
import gpu_util as G
import model_util as MU
from ipavsr_b200 import layers as L
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.utils import evaluate as EV

pytestmark = pytest.mark.gpu

GE = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'evaluate.npz'))


def _run_kernel(probs, mask, y, ldp=None):
    N, T, C = probs.shape
    ldp = ldp or C
    p = np.zeros((N * T, ldp), 'float32')
    p[:, :C] = probs.reshape(N * T, C)
    p[:, C:] = 9.0                                  # padding beyond C must never win
    dp = G.dev(p)
    dm = torch.from_numpy(mask).cuda() if mask is not None else None
    dy = torch.from_numpy(y).cuda()
    pred = torch.full((N,), -1, dtype=torch.int32, device='cuda')
    conf = torch.zeros(C, C, dtype=torch.int32, device='cuda')
    corr = torch.zeros(1, dtype=torch.int32, device='cuda')
    G.call('ipavsr_vote_eval', dp.data_ptr(), ldp, dm.data_ptr() if dm is not None else None, dy.data_ptr(), N, T, C,
           pred.data_ptr(), conf.data_ptr(), corr.data_ptr(), G.stream())
    return pred.cpu().numpy(), conf.cpu().numpy(), int(corr.item())


def test_vote_kernel_matches_reference_golden():
    pred, conf, ok = _run_kernel(GE['vote_probs'], GE['vote_mask'], GE['vote_y'])
    np.testing.assert_array_equal(conf, GE['vote_conf'])
    assert ok / float(len(pred)) == float(GE['vote_rate'])
    assert pred[2] == 1 and pred[3] == 2            # tie rules of np.argmax at both levels
    so = GE['seq_probs']
    pred, conf, ok = _run_kernel(so[:, None, :], None, GE['seq_y'])
    np.testing.assert_array_equal(conf, GE['seq_conf'])
    assert ok / float(len(pred)) == float(GE['seq_rate'])


@pytest.mark.parametrize('N,T,C,ldp', [(1, 1, 2, 2), (513, 40, 26, 26), (64, 33, 10, 12), (9, 70, 300, 304), (2000, 5, 3, 4)])
def test_vote_kernel_matches_oracle(N, T, C, ldp):
    rng = np.random.default_rng(N + T + C)
    # few distinct values -> many exact ties, in frames and in votes
    probs = rng.integers(0, 4, size=(N, T, C)).astype('float32') / 4
    lens = rng.integers(0, T + 1, size=N)            # zero-length utterances vote for class 0 (np.argmax of zeros)
    lens[0] = T
    mask = (np.arange(T)[None, :] < lens[:, None]).astype('uint8')
    y = rng.integers(0, C, size=N).astype('uint8') if C <= 256 else rng.integers(0, 256, size=N).astype('uint8')
    pred, conf, ok = _run_kernel(probs, mask, y, ldp)
    want = OE.vote_predictions(probs, mask)
    np.testing.assert_array_equal(pred, want)
    np.testing.assert_array_equal(conf, OE.confusion(want, y, C))
    assert ok == int((want == y).sum())


def test_python_api_on_a_compiled_function():
    """evaluate_model2 (two-stream argument order) on a real network: same numbers as the reference algorithm applied to
    the host copy of the same outputs; the probabilities never leave the device."""
    rng = np.random.default_rng(11)
    spec = MU.build('adenet_v2', rng, fusiontype='concat')
    net = spec['net']
    ins = MU.input_layers(net)
    N, Tm = 19, 13
    feed, mask, lens = MU.make_feed(rng, N, Tm, spec['dims'])
    y = rng.integers(0, 7, size=N).astype('uint8')
    window = T.iscalar('theta')
    val_fn = function([ins['input'].input_var, ins['mask'].input_var, ins['dct'].input_var, window],
                      L.get_output(net, deterministic=True))
    host = val_fn(feed[0], mask, feed[1], 3)
    dev = val_fn(feed[0], mask, feed[1], 3, device_output=True)
    assert isinstance(dev, torch.Tensor) and dev.is_cuda and tuple(dev.shape) == host.shape
    np.testing.assert_array_equal(dev.cpu().numpy(), host)
    rate, conf = EV.evaluate_model2_2stream(feed[0], y, mask, feed[1], 3, val_fn)
    want_rate, want_conf = OE.evaluate_vote(host, y, mask)
    assert rate == want_rate
    np.testing.assert_array_equal(conf, want_conf)
    assert conf.dtype == np.zeros(1, dtype='int').dtype
