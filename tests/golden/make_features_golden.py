"""Generates tests/golden/features.npz by running the REFERENCE's own `utils/preprocessing.py` functions of SURVEY §8f
rank 3 (zigzag, compute_dct_features, reorder_data, force_align, multistream_force_align; imported from /root/reference,
read-only) on seeded inputs.  Run in the build container only:

    python tests/golden/make_features_golden.py

Shims (nothing under /root/reference is modified or copied): `scipy.misc.imresize` and `numpy.matlib` as in
make_golden.py; `xrange` (Python 2 builtin used at utils/preprocessing.py:431) = range; `scipy.fft.dct` is what the
module's `from scipy import fftpack as fft` era call resolves to (same pocketfft float32 transform).
"""
import builtins
import os
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings('ignore')
import scipy.misc                                       # noqa: E402

if not hasattr(scipy.misc, 'imresize'):
    scipy.misc.imresize = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError('shim'))
try:
    import numpy.matlib                                 # noqa: F401
except Exception:                                       # pragma: no cover
    m = types.ModuleType('numpy.matlib')
    m.repmat = lambda a, r, c: np.tile(np.atleast_2d(a), (r, c))
    sys.modules['numpy.matlib'] = m
builtins.xrange = range

sys.path.insert(0, '/root/reference')
from utils import preprocessing as ref                  # noqa: E402


def zigzag_or_error(rows, cols):
    try:
        return ref.zigzag(np.arange(rows * cols).reshape(rows, cols))
    except IndexError:
        return np.array([-1])


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    shapes = [(r, c) for r in range(1, 9) for c in range(1, 9)] + [(30, 40), (26, 44), (44, 26)]
    out['zigzag_shapes'] = np.array(shapes)
    for r, c in shapes:
        out['zigzag_%d_%d' % (r, c)] = zigzag_or_error(r, c)
    # compute_dct_features on 6 x 8 images, every method; the columns get distinct scales so that the selection
    # methods' argsort has no near-ties
    frames, shape = 37, (6, 8)
    X = rng.normal(0.5, 1.0, size=(frames, 48)).astype(np.float32)
    X += (np.cos(np.outer(np.arange(frames), np.arange(48)) * 0.37) * np.linspace(0.2, 3.0, 48)).astype(np.float32)
    out['dct_X'] = X
    out['dct_shape'] = np.array(shape)
    for method in ('zigzag', 'variance', 'rel_variance', 'energy'):
        out['dct_' + method] = ref.compute_dct_features(X.copy(), shape, 10, method)
    out['dct_zigzag_30'] = ref.compute_dct_features(X.copy(), shape, 30, 'zigzag')
    # reorder_data both ways
    R = rng.normal(size=(5, 12)).astype(np.float32)
    out['reorder_X'] = R
    out['reorder_f2c'] = ref.reorder_data(R.copy(), (3, 4), 'f', 'c')
    out['reorder_c2f'] = ref.reorder_data(R.copy(), (3, 4), 'c', 'f')
    out['reorder_f2f'] = ref.reorder_data(R.copy(), (3, 4), 'f', 'f')
    # force_align: stream 2's fill frame is indexed with stream 1's length (:652), so the last utterance must not be
    # one where stream 1 is longer (the reference raises IndexError there)
    l1 = np.array([5, 7, 4, 6, 3]); l2 = np.array([7, 4, 4, 5, 6])
    a = rng.normal(size=(int(l1.sum()), 8)).astype(np.float32)
    b = rng.normal(size=(int(l2.sum()), 6)).astype(np.float32)
    ta = np.repeat(np.arange(5), l1).astype(np.uint8); tb = np.repeat(np.arange(5), l2).astype(np.uint8)
    out['fa_l1'], out['fa_l2'], out['fa_a'], out['fa_b'], out['fa_ta'], out['fa_tb'] = l1.copy(), l2.copy(), a, b, ta, tb
    (na, nta, nl1), (nb, ntb, nl2) = ref.force_align((a, ta, l1.copy()), (b, tb, l2.copy()))
    out['fa_out_a'], out['fa_out_ta'], out['fa_out_l1'] = na, nta, nl1
    out['fa_out_b'], out['fa_out_tb'], out['fa_out_l2'] = nb, ntb, nl2
    # multistream_force_align, three streams
    l3 = np.array([6, 6, 9, 2, 4])
    c = rng.normal(size=(int(l3.sum()), 4)).astype(np.float32)
    tc = np.repeat(np.arange(5), l3).astype(np.uint8)
    out['ms_l3'], out['ms_c'], out['ms_tc'] = l3.copy(), c, tc
    res = ref.multistream_force_align([(a, ta, l1.copy()), (b, tb, l2.copy()), (c, tc, l3.copy())])
    for j, (x, t, l) in enumerate(res):
        out['ms_out_x%d' % j], out['ms_out_t%d' % j], out['ms_out_l%d' % j] = x, t, np.asarray(l)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'features.npz'), **out)
    print('wrote features.npz with %d arrays' % len(out))


if __name__ == '__main__':
    main()
