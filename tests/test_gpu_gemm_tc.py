"""tcgen05 GEMM modes (TF32 single pass, 3xTF32 fp32-parity) against a float64 product, all operand layouts."""
import numpy as np
import pytest
import torch

from oracle import ops
import gpu_util as G

pytestmark = pytest.mark.gpu

SHAPES = [(128, 256, 64), (1040, 2000, 1200), (300, 1000, 152), (2048, 52, 500), (1200, 2000, 4096),
          (252, 1000, 1040), (1040, 250, 1000), (129, 130, 200), (4096, 2000, 1200)]


def _run(mode, ta, tb, M, N, K, act=0, acc=0, use_bias=False, seed=0, scale_a=1.0, scale_b=1.0):
    rng = np.random.default_rng(seed + M + N + K)
    lda = ((M if ta else K) + 3) // 4 * 4
    ldb = ((K if tb else N) + 3) // 4 * 4
    ldc = (N + 3) // 4 * 4
    A = np.zeros((K if ta else M, lda), 'float32')
    B = np.zeros((N if tb else K, ldb), 'float32')
    A[:, :(M if ta else K)] = rng.normal(size=(K if ta else M, M if ta else K)) * scale_a
    B[:, :(K if tb else N)] = rng.normal(size=(N if tb else K, K if tb else N)) * scale_b
    bias = rng.normal(size=(N,)).astype('float32')
    C0 = np.zeros((M, ldc), 'float32')
    C0[:, :N] = rng.normal(size=(M, N))
    Aop = (A[:, :M].T if ta else A[:, :K]).astype(np.float64)
    Bop = (B[:, :K].T if tb else B[:, :N]).astype(np.float64)
    ref = Aop @ Bop + (C0[:, :N] if acc else 0) + (bias if use_bias else 0)
    want = ops.act_fwd(ref, act)
    dA, dB, db, dC = G.dev(A), G.dev(B), G.dev(bias), G.dev(C0)
    need = G.lib().ipavsr_gemm_workspace_bytes(mode, ta, tb, M, N, K)
    ws = G.zeros((max(need // 4, 4),))
    G.call('ipavsr_gemm', mode, ta, tb, M, N, K, dA.data_ptr(), lda, dB.data_ptr(), ldb, dC.data_ptr(), ldc,
           db.data_ptr() if use_bias else None, act, acc, ws.data_ptr(), int(need), G.stream())
    got = G.host(dC)
    if ldc > N:
        np.testing.assert_array_equal(got[:, N:], C0[:, N:])          # padding columns untouched
    scale = np.abs(Aop).max() * np.abs(Bop).max() * np.sqrt(K)
    return np.abs(got[:, :N] - want).max() / max(scale, 1e-30) if act == 0 else G.relerr(got[:, :N], want)


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('ta,tb', [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_tf32x3_matches_fp32_parity(M, N, K, ta, tb):
    err = _run(1, ta, tb, M, N, K)
    assert err < 1.5e-6, err          # 3xTF32: fp32-class accuracy (plain TF32 is ~2e-4 in the same units)


@pytest.mark.parametrize('M,N,K', SHAPES[:5])
@pytest.mark.parametrize('ta,tb', [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_tf32_single_pass(M, N, K, ta, tb):
    err = _run(2, ta, tb, M, N, K)
    assert err < 1e-3, err            # stated tolerance of the TF32 mode (10-bit mantissa operands)


@pytest.mark.parametrize('act,acc,use_bias', [(1, 0, True), (2, 1, True), (0, 1, False), (3, 0, True)])
def test_epilogue_variants(act, acc, use_bias):
    for mode, tol in ((1, 5e-4), (2, 1e-1)):
        err = _run(mode, 0, 0, 1040, 2000, 1200, act, acc, use_bias)
        assert err < tol, (mode, err)


def test_three_term_split_beats_single_pass():
    e3 = _run(1, 0, 0, 1040, 1000, 2000)
    e1 = _run(2, 0, 0, 1040, 1000, 2000)
    assert e3 < e1 / 50, (e3, e1)


# ---- fp16 three-product mode (IPAVSR_GEMM_F16X3 = 4): same accuracy class as 3xTF32 -----------------------------
@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('ta,tb', [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_f16x3_matches_fp32_parity(M, N, K, ta, tb):
    err = _run(4, ta, tb, M, N, K)
    assert err < 1.5e-6, err


# ---- persistent kernel (csrc/gemm_f16p.cu): products with >= 2 work units per CTA pair and no k-split -------------------
# All three partial products accumulate in ONE fp32 TMEM accumulator (three truncating adds per k-step instead of one): its
# stated tolerance is 4e-6 of |A|max |B|max sqrt(K) (measured 1.3e-6 at K = 1200, 2.7e-6 at K = 4096) against 1.5e-6 for
# the tile-per-pair kernel with its separate cross-term accumulator.
@pytest.mark.parametrize('ta,tb,M,N,K', [(0, 0, 13325, 2000, 1200), (0, 1, 13325, 2000, 1000), (1, 0, 9728, 2000, 1024),
                                         (1, 1, 9728, 2000, 1024), (0, 0, 9700, 2040, 4000)])
def test_f16x3_persistent_kernel(ta, tb, M, N, K):
    before = G.lib().ipavsr_debug_gemm_persistent_launches()
    err = _run(4, ta, tb, M, N, K)
    assert err < 4e-6, err
    assert G.lib().ipavsr_debug_gemm_persistent_launches() == before + 1        # the persistent kernel ran this product


@pytest.mark.parametrize('act,acc,use_bias', [(1, 0, True), (2, 1, True), (0, 1, False), (3, 0, True)])
def test_f16x3_persistent_epilogue_variants(act, acc, use_bias):
    err = _run(4, 0, 0, 13325, 2000, 1200, act, acc, use_bias)
    assert err < 5e-4, err


@pytest.mark.parametrize('sa,sb', [(1e-7, 1.0), (3e4, 1e-3), (1e-12, 1e-9), (1e6, 1e5)])
def test_f16x3_scale_invariance(sa, sb):
    """Per-tensor power-of-two scales: operand magnitudes far outside the fp16 range keep fp32-class accuracy."""
    for ta, tb in ((0, 0), (1, 0)):
        err = _run(4, ta, tb, 1040, 1000, 1200, scale_a=sa, scale_b=sb)
        assert err < 1.5e-6, (ta, tb, err)


@pytest.mark.parametrize('act,acc,use_bias', [(1, 0, True), (2, 1, True), (0, 1, False), (3, 0, True)])
def test_f16x3_epilogue_variants(act, acc, use_bias):
    err = _run(4, 0, 0, 1040, 2000, 1200, act, acc, use_bias)
    assert err < 5e-4, err


def test_f16x3_presplit_api_and_amax():
    """ipavsr_f16_split (+ producer-side amax) -> ipavsr_gemm_f16x3, with an outlier-heavy operand."""
    rng = np.random.default_rng(7)
    M, N, K = 1040, 504, 1208
    A = (rng.normal(size=(M, K)) * np.exp(rng.normal(size=(M, K)) * 3)).astype('float32')      # ~2^30 dynamic range
    B = rng.normal(size=(K, N)).astype('float32') * 1e-4
    dA, dB = G.dev(A), G.dev(B)
    ah, al = G.zeros((M, K), torch.float16), G.zeros((M, K), torch.float16)
    bh, bl = G.zeros((K, N), torch.float16), G.zeros((K, N), torch.float16)
    amax, exps = G.zeros((4,)), G.zeros((4,), torch.int32)
    G.call('ipavsr_f16_split', dA.data_ptr(), K, M, K, ah.data_ptr(), al.data_ptr(), K, amax.data_ptr(), exps.data_ptr(), 0,
           G.stream())
    G.call('ipavsr_amax', dB.data_ptr(), N, K, N, amax.data_ptr() + 4, G.stream())
    G.call('ipavsr_f16_split', dB.data_ptr(), N, K, N, bh.data_ptr(), bl.data_ptr(), N, amax.data_ptr() + 4,
           exps.data_ptr() + 4, 1, G.stream())
    am = G.host(amax)
    assert am[0] == np.abs(A).max() and am[1] == np.abs(B).max()
    e = G.host(exps)
    assert 2.0 ** 14 <= am[0] * 2.0 ** int(e[0]) < 2.0 ** 15 and 2.0 ** 14 <= am[1] * 2.0 ** int(e[1]) < 2.0 ** 15
    # the pair reproduces the scaled value to ~2^-22
    rec = (G.host(ah).astype(np.float64) + G.host(al).astype(np.float64) / G.F16_LO_SCALE) / 2.0 ** int(e[0])
    # (the unscaled lo half sits in the fp16 subnormals below 2^-18 of the tensor's maximum: an absolute floor there)
    big = np.abs(A) > np.abs(A).max() * 2.0 ** -17
    assert (np.abs(rec - A)[big] <= np.abs(A)[big] * 2.0 ** -21).all()
    assert (np.abs(rec - A) <= np.maximum(np.abs(A) * 2.0 ** -21, np.abs(A).max() * 2.0 ** -38)).all()
    dC = G.zeros((M, N))
    G.call('ipavsr_gemm_f16x3', 0, 0, M, N, K, ah.data_ptr(), al.data_ptr(), K, exps.data_ptr(), bh.data_ptr(),
           bl.data_ptr(), N, exps.data_ptr() + 4, dC.data_ptr(), N, None, 0, 0, amax.data_ptr() + 8, None, None, 0, G.stream())
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = G.host(dC)
    # error relative to the natural scale of each dot product (sum of |a||b|)
    nat = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    assert (np.abs(got - ref) / nat).max() < 8e-6      # dominated by the fp32 accumulation of 76 MMAs, as in 3xTF32
    assert G.host(amax)[2] == np.abs(got).max()


def test_f16x3_beats_single_pass_tf32():
    e3 = _run(4, 0, 0, 1040, 1000, 2000)
    e1 = _run(2, 0, 0, 1040, 1000, 2000)
    assert e3 < e1 / 50, (e3, e1)


def test_f16x3_epilogue_emits_static_scale_split():
    """C_hi / C_lo written by the epilogue (sigmoid output, static exponent 14) == ipavsr_f16_split of C with amax = 1."""
    rng = np.random.default_rng(11)
    M, N, K = 1040, 1000, 1200
    A = rng.normal(size=(M, K)).astype('float32')
    B = (rng.normal(size=(K, N)) / 30).astype('float32')
    dA, dB = G.dev(A), G.dev(B)
    ah, al = G.zeros((M, K), torch.float16), G.zeros((M, K), torch.float16)
    bh, bl = G.zeros((K, N), torch.float16), G.zeros((K, N), torch.float16)
    sc = G.zeros((4,))
    G.call('ipavsr_f16_split', dA.data_ptr(), K, M, K, ah.data_ptr(), al.data_ptr(), K, sc.data_ptr(), sc.data_ptr() + 4, 0,
           G.stream())
    G.call('ipavsr_f16_split', dB.data_ptr(), N, K, N, bh.data_ptr(), bl.data_ptr(), N, sc.data_ptr() + 8, sc.data_ptr() + 12,
           0, G.stream())
    dC = G.zeros((M, N))
    ch, cl = G.zeros((M, N), torch.float16), G.zeros((M, N), torch.float16)
    G.call('ipavsr_gemm_f16x3', 0, 0, M, N, K, ah.data_ptr(), al.data_ptr(), K, sc.data_ptr() + 4, bh.data_ptr(),
           bl.data_ptr(), N, sc.data_ptr() + 12, dC.data_ptr(), N, None, 1, 0, None, ch.data_ptr(), cl.data_ptr(), 14,
           G.stream())
    wh, wl = G.zeros((M, N), torch.float16), G.zeros((M, N), torch.float16)
    one = G.dev(np.array([1.0, 0.0], 'float32'))
    G.call('ipavsr_f16_split', dC.data_ptr(), N, M, N, wh.data_ptr(), wl.data_ptr(), N, one.data_ptr(), one.data_ptr() + 4, 1,
           G.stream())
    assert int(G.host(one).view(np.int32)[1]) == 14
    np.testing.assert_array_equal(G.host(ch).view(np.uint16), G.host(wh).view(np.uint16))
    np.testing.assert_array_equal(G.host(cl).view(np.uint16), G.host(wl).view(np.uint16))
