"""Generates tests/golden/preprocessing.npz by running the REFERENCE's own `utils/preprocessing.py`
(imported from /root/reference, read-only) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference module imports `scipy.misc.imresize` (removed from SciPy) and `numpy.matlib` at import time;
both are shimmed here — `imresize` is never called by the six functions on the hot path, `numpy.matlib.repmat`
is the real NumPy implementation.  Nothing under /root/reference is modified or copied.
"""
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

sys.path.insert(0, '/root/reference')
from utils import preprocessing as ref                  # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    # utterance lengths are passed as a list of Python ints: under NumPy >= 2 (NEP 50) `float32_array / np.int64`
    # promotes to float64, whereas the reference's NumPy 1.x era value-based casting keeps float32; a Python int
    # divisor reproduces the era's float32 arithmetic in `sequencewise_mean_image_subtraction` (:273)
    lens_list = [12, 9, 20, 15, 11]
    lens = np.array(lens_list, dtype=np.int64)
    frames = int(lens.sum())
    X = rng.normal(3.0, 2.0, size=(frames, 48)).astype(np.float32)
    out['lens'] = lens
    out['X'] = X
    out['normalize_input'] = ref.normalize_input(X.copy())
    n, mean, std = ref.featurewise_normalize_sequence(X.copy())
    out['featurewise_norm'], out['featurewise_mean'], out['featurewise_std'] = n, mean, std
    out['seq_mean_sub'] = ref.sequencewise_mean_image_subtraction(X.copy(), lens_list)
    out['diff_images'] = ref.compute_diff_images(X.copy(), lens_list)
    Xd = rng.normal(size=(frames, 30)).astype(np.float32)
    out['Xd'] = Xd
    out['concat_deltas_w9'] = ref.concat_first_second_deltas(Xd.copy(), lens_list, 9)
    out['concat_deltas_w5'] = ref.concat_first_second_deltas(Xd.copy(), lens_list, 5)
    a = np.array([[1, 1, 1, 1, 1, 1, 1, 1, 10], [2, 2, 2, 2, 2, 2, 2, 2, 20],
                  [3, 3, 3, 3, 3, 3, 3, 3, 30], [4, 4, 4, 4, 4, 4, 4, 4, 40]])       # utils/preprocessing.py:12
    out['test_delta_in'] = a
    out['test_delta_out'] = ref.deltas(a, 9)
    q = np.array([[5, 1, 2, 3, 4, 5, 6, 7, 8, 9]])
    out['leftpad_in'] = q
    out['leftpad_out'] = ref.deltas(q, 9)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'preprocessing.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == '__main__':
    main()
