"""Property tests (hypothesis) of the host-side logic and of the oracle: invariants the domain offers that do not depend
on sizes ("hypothesis property tests over (N, T, F, theta, lens)").  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from ipavsr_b200.engine import _PackPlan
from oracle import ops, preprocessing as OP


@st.composite
def _lens(draw):
    T = draw(st.integers(1, 12))
    N = draw(st.integers(1, 40))
    lens = draw(st.lists(st.integers(0, T), min_size=N, max_size=N))
    return T, np.asarray(lens, dtype=np.int64)


@settings(max_examples=60, deadline=None)
@given(_lens(), st.integers(0, 2 ** 31 - 1))
def test_pack_plan_is_a_consistent_permutation(tl, seed):
    """pack / unpack / perm / unperm / order / inv of the packed, length-sorted execution are mutually consistent for any
    lengths (zero-length utterances and all-full batches included)."""
    T, lens = tl
    N = len(lens)
    plan = _PackPlan(lens, T)
    t = dict(plan.tables)
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(N * T, 2))
    mask = (np.arange(T)[None, :] < lens[:, None])
    x[~mask.reshape(-1)] = 0                       # the reference zero-pads past the length (utils/datagen.py:138-139)
    assert plan.M == int(lens.sum())
    assert (np.diff(plan.lens_sorted) <= 0).all()
    assert sorted(t['order'].tolist()) == list(range(N))
    packed = np.where(t['pack'][:, None] >= 0, x[np.maximum(t['pack'], 0)], 0.0)
    srt = x.reshape(N, T, 2)[plan.order_host].reshape(N * T, 2)
    np.testing.assert_array_equal(packed[t['unpack']], srt)
    np.testing.assert_array_equal(srt[t['valid']], packed[:-1])
    np.testing.assert_array_equal(srt[t['unperm']], x)
    np.testing.assert_array_equal(x[t['perm']], srt)
    np.testing.assert_array_equal(t['mask'].view(np.uint8)[:N * T].reshape(N, T), mask[plan.order_host])
    # rows a per-step recurrent GEMM has to compute: utterances longer than t are a prefix of the sorted batch
    for tt in range(T):
        k = int(plan.active_rows[tt])
        assert (plan.lens_sorted[:k] > tt).all() and (plan.lens_sorted[k:] <= tt).all()
    assert (plan.offsets_host == np.concatenate([[0], np.cumsum(plan.lens_sorted)])).all()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 4), st.integers(1, 9), st.integers(1, 5), st.integers(1, 9), st.integers(0, 2 ** 31 - 1))
def test_delta_layer_is_linear_and_theta1_is_a_central_difference(N, T, F, theta, seed):
    """DeltaLayer (utils/signal.py:59-80) is a linear operator; with theta = 1 its delta block is (x[t+1] - x[t-1]) / 2 with
    edge replication; its backward is the adjoint of its forward."""
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(N, T, F)).astype('float32')
    y = rng.normal(size=(N, T, F)).astype('float32')
    a = np.float32(rng.normal())
    f = lambda v: ops.delta_fwd(v, theta).astype(np.float64)
    lhs = f((x + a * y).astype('float32'))
    rhs = f(x) + np.float64(a) * f(y)
    assert np.abs(lhs - rhs).max() <= 1e-4 * max(1.0, np.abs(rhs).max())
    d1 = ops.delta_fwd(x, 1)[..., F:2 * F].astype(np.float64)
    xp = np.concatenate([x[:, :1], x, x[:, -1:]], axis=1).astype(np.float64)
    np.testing.assert_allclose(d1, (xp[:, 2:] - xp[:, :-2]) / 2.0, atol=1e-5)
    # the backward (ops.delta_bwd) is the adjoint of the forward: <D x, g> = <x, D^T g>
    g = rng.normal(size=(N, T, 3 * F))
    gx = ops.delta_bwd(g, theta, np.float64)
    lhs = float((f(x) * g).sum())
    rhs = float((x.astype(np.float64) * gx).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs), float(np.abs(g).sum()))


@settings(max_examples=40, deadline=None)
@given(_lens(), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_diff_images_telescoping_and_mean_subtraction_idempotence(tl, D, seed):
    """compute_diff_images (utils/preprocessing.py:506-517): the differences of an utterance telescope back to last - first,
    and its first row duplicates the second; sequencewise_mean_image_subtraction (:260-277) is idempotent up to rounding."""
    T, lens = tl
    lens = np.maximum(lens, 2)
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(int(lens.sum()), D)).astype('float32')
    d = OP.compute_diff_images(x, lens)
    o = 0
    for n in lens:
        seg, dd = x[o:o + n].astype(np.float64), d[o:o + n].astype(np.float64)
        np.testing.assert_array_equal(dd[0], dd[1])
        np.testing.assert_allclose(dd[1:].sum(0), seg[-1] - seg[0], atol=1e-4)
        o += n
    m1 = OP.sequencewise_mean_image_subtraction(x, lens)
    m2 = OP.sequencewise_mean_image_subtraction(m1, lens)
    np.testing.assert_allclose(m2, m1, atol=1e-5)
