"""CPU: the oracle port against the known-answer / self-consistency tests the reference's int8 libraries
carry for this path: gemmology test_quantize.cpp:56-62, test_multiply.cpp:152-174
(Shift PrepareA == +127), :326-386 (PrepareBias vs slow int), :224-324,414-425 (ShiftInt shapes)."""
import numpy as np
import pytest

from oracle import slimt_oracle as so
import sb_testutil as util

CORNERS = np.array([-32769., -32768., -32767., -129., -128., -127., -1., 0., 1., 126., 127., 128., 129., 32766., 32768.,
                    32769., -1.9, -1.5, -1.1, -1., -0.9, -0.5, -0.1, 0.0, 0.1, 0.5, 0.9, 1.0, 1.1, 1.5, 1.9, 16056.8, 2.5],
                   dtype=np.float32)


def quantize_ref(x, mult):
    """QuantizeRef of gemmology/test/test_quantize.cpp:9-16 (roundf, clamp to +-127)."""
    v = np.float32(x) * np.float32(mult)
    r = np.sign(v) * np.floor(np.abs(v) + np.float32(0.5))
    return np.clip(r, -127, 127).astype(np.int8)


@pytest.mark.parametrize("mult", [1.0, 32.0, -1.0, -0.49])
def test_quantize_corner_values(mult):
    q = so.quantize(CORNERS, mult)
    ref = quantize_ref(CORNERS, mult)
    prod = CORNERS * np.float32(mult)
    for i in range(len(CORNERS)):
        if q[i] == ref[i]:
            continue
        # IsOff(): exact .5 cases may round either way (the SIMD path rounds to even, roundf away from zero)
        off_t, off_r = abs(float(q[i]) - prod[i]), abs(float(ref[i]) - prod[i])
        assert abs(int(q[i]) - int(ref[i])) <= 1 and 0.49 < off_t < 0.51 and 0.49 < off_r < 0.51, (CORNERS[i], mult, q[i], ref[i])


def test_quantize_ties_round_to_even_and_overflow():
    x = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 126.5, 127.5, 1e30, -1e30, np.nan, 3e9, -3e9], dtype=np.float32)
    q = so.quantize(x, 1.0)
    # cvtps2dq: RNE; NaN and |t| >= 2^31 give INT_MIN which the clamp turns into -127
    assert q.tolist() == [0, 2, 2, 0, -2, -2, 126, 127, -127, -127, -127, -127, -127]


def test_shift_prepare_a_is_plus_127():
    rng = np.random.RandomState(3)
    x = rng.uniform(-2, 2, size=(8, 256)).astype(np.float32)
    qa = so.quantize(x, 64.0)
    u8 = qa.astype(np.int16) + 127
    assert u8.min() >= 0 and u8.max() <= 254
    assert np.array_equal(qa, quantize_ref(x, 64.0)) or np.abs(qa.astype(int) - quantize_ref(x, 64.0).astype(int)).max() <= 1


@pytest.mark.parametrize("shape", [(256, 256), (2048, 256), (256, 1536)])
def test_prepare_bias_against_slow_int(shape):
    K, N = shape
    rng = np.random.RandomState(K + N)
    Bt = rng.randint(-127, 128, size=(N, K)).astype(np.int8)
    bias = rng.standard_normal(N).astype(np.float32)
    aq, bq = 127.0 / 2.0, 127.0 / 0.3
    pb, colsum = so.prepare_bias(Bt, bias, aq, bq)
    slow = Bt.astype(np.int64).sum(axis=1)
    assert np.array_equal(colsum.astype(np.int64), slow)
    m = np.float32(-1.0) * (np.float32(127.0 / aq) * np.float32(127.0 / bq)) / np.float32(127.0)
    assert np.allclose(pb, slow.astype(np.float32) * m + bias, rtol=0, atol=1e-4)


# (A_rows, width, B_cols, float_tolerance, MSE_float_tolerance) of gemmology/test/test_multiply.cpp:414-425
SHIFT_INT_CASES = [(8, 256, 256, 0.54, 0.17), (8, 2048, 256, 1.66, 0.46), (320, 256, 256, 0.64, 0.16),
                   (472, 256, 256, 0.62, 0.17), (248, 256, 256, 0.64, 0.16), (200, 256, 256, 0.74, 0.17)]


@pytest.mark.parametrize("case", SHIFT_INT_CASES)
def test_multiply_shift_int(case):
    """TestMultiplyShiftInt (test_multiply.cpp:224-324): alpha = 2, operands ~ U(-1, 1); the shifted int8 product
    must equal the slow integer reference (here: exactly) and track the float product within the reference's
    own tolerances."""
    M, K, N, float_tol, mse_tol = case
    rng = np.random.RandomState(M * 7 + K)
    A = rng.uniform(-1, 1, size=(M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, size=(N, K)).astype(np.float32)
    bias = rng.uniform(-1, 1, size=N).astype(np.float32)
    aq = bq = 127.0 / 2.0
    Bq = so.quantize(B, bq)
    y, qa, acc = so.affine(A, Bq, bias, aq, bq, want=True)
    exact_int = (qa.astype(np.int64) + 127) @ Bq.astype(np.int64).T
    assert np.array_equal(acc.astype(np.int64), exact_int)  # float64-BLAS shortcut == exact integer product
    assert np.array_equal(so.gemm_shifted_c(qa, Bq), acc)    # == the scalar C loop
    ref = A.astype(np.float64) @ B.astype(np.float64).T + bias
    assert np.abs(y - ref).max() <= float_tol
    assert np.sqrt(np.mean((y - ref) ** 2)) <= mse_tol


def test_saturation_count_and_maddubs_mode():
    """All-positive large operands overflow int16 pair sums on non-VNNI x86 (appendix A.4)."""
    qa = np.full((4, 64), 127, dtype=np.int8)
    Bt = np.full((8, 64), 127, dtype=np.int8)
    assert so.saturation_count(qa, Bt) == 4 * 8 * 32
    exact = so.gemm_shifted(qa, Bt, exact=True)
    sat = so.gemm_shifted(qa, Bt, exact=False)
    assert (exact == 64 * 254 * 127).all() and (sat == 32 * 32767).all()
    rng = np.random.RandomState(0)
    qa = rng.randint(-40, 40, size=(8, 256)).astype(np.int8)
    Bt = rng.randint(-60, 60, size=(16, 256)).astype(np.int8)
    assert so.saturation_count(qa, Bt) == 0
    assert np.array_equal(so.gemm_shifted(qa, Bt, exact=True), so.gemm_shifted(qa, Bt, exact=False))


def test_argmax_first_maximum_on_ties():
    x = np.array([[1, 5, 5, 2], [0, 0, 0, 0], [-1, -3, -1, -2], [-0.0, 0.0, -5, 0]], dtype=np.float32)
    assert so.argmax_first(x).tolist() == [1, 0, 0, 0]


def test_shortlist_generate_properties():
    from slimt_b200 import synth
    fr, offs, lists = synth.make_shortlist(vocab=2000, frequent=20, best=7, seed=3, spread=100)
    words = np.array([5, 1999, 17, 5, 0], dtype=np.uint32)
    sl = so.shortlist_generate(words, fr, offs, lists, 2000)
    assert len(sl) % 8 == 0 and np.all(np.diff(sl.astype(np.int64)) > 0)
    assert set(range(20)) <= set(sl.tolist())
    for w in (5, 1999, 17, 0):
        assert set(lists[int(offs[w]):int(offs[w + 1])].tolist()) <= set(sl.tolist())
