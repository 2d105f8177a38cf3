"""-m gpu parity tests of SURVEY §8f rank 3 (csrc/features.cu through the C-ABI and the utils.preprocessing mirror):
compute_dct_features, reorder_data, force_align, multistream_force_align against the reference's own outputs
(tests/golden/features.npz) and against the CPU oracle on seeded inputs at sizes it finishes in seconds."""
import os

import numpy as np
import pytest

from oracle import preprocessing as OP

FG = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'features.npz'))
# float32 accumulation over D terms against the float64 oracle / scipy's float32 FFT: stated tolerance, relative to the
# largest coefficient
DCT_TOL = 2e-5


def _pre():
    from ipavsr_b200.utils import preprocessing as P
    return P


@pytest.mark.gpu
def test_dct_features_golden_every_method():
    P = _pre()
    X, shape = FG['dct_X'], tuple(int(v) for v in FG['dct_shape'])
    for key, k, method in (('dct_zigzag', 10, 'zigzag'), ('dct_zigzag_30', 30, 'zigzag'), ('dct_variance', 10, 'variance'),
                           ('dct_rel_variance', 10, 'rel_variance'), ('dct_energy', 10, 'energy')):
        got = P.compute_dct_features(X, shape, k, method)
        want = FG[key]
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.abs(got - want).max() <= DCT_TOL * np.abs(want).max(), key


@pytest.mark.gpu
@pytest.mark.parametrize('frames,shape,k', [(1, (30, 40), 30), (127, (30, 40), 30), (1000, (26, 44), 30), (300, (30, 50), 45),
                                            (257, (7, 9), 62), (4096, (30, 40), 30),
                                            # the tensor-core kernel (csrc/dct_tc.cu: frames >= 1024, K <= 32): ragged last
                                            # frame tile, D = 1144 (k tail inside a 32-float block), K < 32 and K = 32
                                            (5000, (26, 44), 30), (2000, (30, 40), 17), (1500, (30, 40), 32)])
def test_dct_zigzag_vs_oracle(frames, shape, k):
    P = _pre()
    rng = np.random.default_rng(frames + k)
    X = rng.normal(0.3, 1.0, size=(frames, shape[0] * shape[1])).astype(np.float32)
    got = P.compute_dct_features(X, shape, k)
    want = OP.compute_dct_features(X, shape, k)
    assert got.shape == want.shape == (frames, min(k, shape[0] * shape[1] - 1))
    assert np.abs(got - want).max() <= DCT_TOL * np.abs(want).max()


@pytest.mark.gpu
def test_dct_is_orthonormal_and_linear():
    """Size-independent properties: the full transform preserves every frame's energy (Parseval) and is linear."""
    import torch
    P = _pre()
    rng = np.random.default_rng(5)
    D, frames = 1200, 2048
    X = rng.normal(size=(frames, D)).astype(np.float32)
    Y = rng.normal(size=(frames, D)).astype(np.float32)
    full = lambda a: P._dct_project(torch.from_numpy(a).cuda(), np.arange(D)).cpu().numpy()      # noqa: E731
    fx, fy, fxy = full(X), full(Y), full(X + 2 * Y)
    np.testing.assert_allclose((fx.astype(np.float64) ** 2).sum(1), (X.astype(np.float64) ** 2).sum(1), rtol=1e-5)
    assert np.abs(fxy - (fx + 2 * fy)).max() <= 1e-4 * np.abs(fxy).max()


@pytest.mark.gpu
def test_zigzag_host_index_walk_matches_reference():
    P = _pre()
    for r, c in FG['zigzag_shapes']:
        want = FG['zigzag_%d_%d' % (r, c)]
        if want[0] == -1:
            with pytest.raises(IndexError):
                P.zigzag_order(int(r), int(c))
        else:
            np.testing.assert_array_equal(P.zigzag(np.arange(r * c).reshape(r, c)), want)
    np.testing.assert_array_equal(P.zigzag(P.fill_zigzag((3, 4))), np.arange(1, 13))


@pytest.mark.gpu
def test_reorder_data_golden_and_round_trip():
    P = _pre()
    for key, a, b in (('reorder_f2c', 'f', 'c'), ('reorder_c2f', 'c', 'f'), ('reorder_f2f', 'f', 'f')):
        np.testing.assert_array_equal(P.reorder_data(FG['reorder_X'], (3, 4), a, b), FG[key])
    rng = np.random.default_rng(3)
    for frames, shape in ((1, (30, 40)), (777, (26, 44)), (5000, (30, 40)), (3, (120, 110))):
        X = rng.normal(size=(frames, shape[0] * shape[1])).astype(np.float32)
        c = P.reorder_data(X, shape)
        np.testing.assert_array_equal(c, OP.reorder_data(X, shape))
        np.testing.assert_array_equal(P.reorder_data(c, shape, 'c', 'f'), X)       # round trip is the identity


@pytest.mark.gpu
def test_force_align_golden():
    P = _pre()
    (na, nta, nl1), (nb, ntb, nl2) = P.force_align((FG['fa_a'], FG['fa_ta'], FG['fa_l1'].copy()),
                                                   (FG['fa_b'], FG['fa_tb'], FG['fa_l2'].copy()))
    for got, key in ((na, 'fa_out_a'), (nta, 'fa_out_ta'), (nl1, 'fa_out_l1'), (nb, 'fa_out_b'), (ntb, 'fa_out_tb'),
                     (nl2, 'fa_out_l2')):
        np.testing.assert_array_equal(got, FG[key])
        assert got.dtype == FG[key].dtype
    # stream 1 longer in the LAST utterance: the reference's x2 fill index (:652) runs off the end -> IndexError
    with pytest.raises(IndexError):
        P.force_align((np.zeros((8, 4), 'float32'), np.zeros(8, 'uint8'), np.array([3, 5])),
                      (np.zeros((5, 4), 'float32'), np.zeros(5, 'uint8'), np.array([3, 2])))


@pytest.mark.gpu
def test_multistream_force_align_golden_and_oracle():
    P = _pre()
    res = P.multistream_force_align([(FG['fa_a'], FG['fa_ta'], FG['fa_l1'].copy()),
                                     (FG['fa_b'], FG['fa_tb'], FG['fa_l2'].copy()),
                                     (FG['ms_c'], FG['ms_tc'], FG['ms_l3'].copy())])
    for j, (x, t, l) in enumerate(res):
        np.testing.assert_array_equal(x, FG['ms_out_x%d' % j])
        np.testing.assert_array_equal(t, FG['ms_out_t%d' % j])
        np.testing.assert_array_equal(l, FG['ms_out_l%d' % j])
    # ragged case at a size with many utterances, odd feature widths (scalar copy path) and 1-frame utterances
    rng = np.random.default_rng(11)
    U = 3000
    lens = [rng.integers(1, 41, size=U) for _ in range(3)]
    xs = [rng.normal(size=(int(l.sum()), F)).astype(np.float32) for l, F in zip(lens, (1200, 90, 37))]
    ts = [np.repeat(np.arange(U) % 26, l).astype(np.uint8) for l in lens]
    got = P.multistream_force_align([(x, t, l.copy()) for x, t, l in zip(xs, ts, lens)])
    want = OP.multistream_force_align([(x, t, l.copy()) for x, t, l in zip(xs, ts, lens)])
    for (gx, gt, gl), (wx, wt, wl) in zip(got, want):
        np.testing.assert_array_equal(gx, wx)
        np.testing.assert_array_equal(gt, wt)
        np.testing.assert_array_equal(gl, wl)
    assert got[0][0].shape[0] == got[1][0].shape[0] == got[2][0].shape[0]          # aligned: equal frame counts
