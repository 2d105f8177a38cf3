"""Generates tests/golden/datagen.npz by running the REFERENCE's own `utils/datagen.py` batch builders (imported from
/root/reference, read-only) on seeded packed datasets.  Run in the build container only:

    python tests/golden/make_datagen_golden.py

`utils/datagen.py` imports `utils.io`, which imports lasagne/theano at module level for its pickling helpers; the
in-memory batch builders exercised here never touch them, so empty stand-in modules are registered for the import.
Nothing under /root/reference is modified or copied.
"""
import os
import sys
import types

import numpy as np

for _name in ('lasagne', 'lasagne.layers', 'theano', 'theano.tensor'):
    sys.modules.setdefault(_name, types.ModuleType(_name))
sys.path.insert(0, '/root/reference')
from utils import datagen as ref                        # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    # packed variable-length storage: (total_frames, F) + per-frame labels + per-utterance lengths
    seqlen = np.array([12, 9, 20, 15, 11, 20, 1, 7, 18, 13, 5], dtype=np.int64)
    U, F = len(seqlen), 10
    X = rng.normal(size=(int(seqlen.sum()), F)).astype('float32')
    y = np.repeat(rng.integers(0, 26, size=U).astype('uint8'), seqlen)
    out['seqlen'], out['X'], out['y'] = seqlen, X, y
    integral = ref.compute_integral_len(seqlen)
    out['integral'] = np.array(integral, dtype=np.int64)
    # gen_seq_batch_from_idx (datagen.py:219-229): an arbitrary index list, max_timesteps above the longest utterance
    idxs = np.array([3, 0, 10, 6, 2, 2], dtype=np.int64)
    out['idxs'] = idxs
    out['seq_batch_T20'] = ref.gen_seq_batch_from_idx(X, idxs, seqlen, integral, 20)
    out['seq_batch_T23'] = ref.gen_seq_batch_from_idx(X, idxs, seqlen, integral, 23)
    # gen_lstm_batch_random (datagen.py:92-153): seeded shuffling, batch size 4 -> batches of 4,4,3 then a new epoch
    np.random.seed(1234)
    g = ref.gen_lstm_batch_random(X, y, seqlen, batchsize=4, shuffle=True)
    for k in range(5):
        xb, yb, mb, ib = next(g)
        out['rand%d_X' % k], out['rand%d_y' % k], out['rand%d_mask' % k] = xb, yb, mb
        out['rand%d_idx' % k] = np.asarray(ib, dtype=np.int64)
    # unshuffled, batch size dividing the number of utterances exactly is not possible with 11; use 11 -> one batch/epoch
    g = ref.gen_lstm_batch_random(X, y, seqlen, batchsize=11, shuffle=False)
    for k in range(2):
        xb, yb, mb, ib = next(g)
        out['seq%d_X' % k], out['seq%d_y' % k], out['seq%d_mask' % k] = xb, yb, mb
        out['seq%d_idx' % k] = np.asarray(list(ib), dtype=np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'datagen.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
