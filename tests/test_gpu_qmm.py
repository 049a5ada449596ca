"""GPU: the qmm:: operator contract through the C ABI (slimt_b200_qmm_affine*), against the oracle.
Integer work (quantized operands, int32 accumulators) and the f32 outputs are all BIT-EXACT."""
import os

import numpy as np
import pytest

import sb_testutil as util
from oracle import slimt_oracle as so

pytestmark = pytest.mark.gpu

SHAPES = util.REFERENCE_GEMM_SHAPES + util.HOTPATH_GEMM_SHAPES + [(1, 256, 256), (129, 256, 264), (300, 512, 512), (1000, 256, 64), (77, 192, 40)]


@pytest.mark.parametrize("shape", SHAPES)
def test_affine_bit_exact(gpu_ctx, shape):
    M, K, N = shape
    x, Bt, bias, aq, bq = util.make_qmm_case(M * 3 + K + N, M, K, N)
    y, qa, acc = gpu_ctx.qmm_affine(x, Bt, bias, aq, bq, debug=True)
    yo, qao, acco = so.affine(x, Bt, bias, aq, bq, want=True)
    assert np.array_equal(qa, qao), "quantized operands differ"
    assert np.array_equal(acc, acco), "int32 accumulators differ"
    assert np.array_equal(y, yo), "f32 outputs differ"


def test_dot_has_no_bias(gpu_ctx):
    x, Bt, _, aq, bq = util.make_qmm_case(5, 40, 256, 256)
    assert np.array_equal(gpu_ctx.qmm_affine(x, Bt, None, aq, bq), so.affine(x, Bt, None, aq, bq))


@pytest.mark.parametrize("n_idx", [8, 512, 4000])
def test_affine_with_select(gpu_ctx, n_idx):
    x, Bt, bias, aq, bq = util.make_qmm_case(n_idx, 24, 256, 32000)
    idx = np.sort(np.random.RandomState(n_idx).choice(32000, n_idx, replace=False)).astype(np.uint32)
    y, qa, acc = gpu_ctx.qmm_affine(x, Bt, bias, aq, bq, idx, debug=True)
    yo, qao, acco = so.affine(x, Bt, bias, aq, bq, indices=idx, want=True)
    assert np.array_equal(qa, qao) and np.array_equal(acc, acco) and np.array_equal(y, yo)


def test_corner_activations(gpu_ctx):
    """test_quantize.cpp corner values plus cvtps2dq overflow / NaN semantics."""
    x = np.zeros((8, 256), dtype=np.float32)
    corners = np.array([-32769., -32768., -32767., -129., -128., -127., -1., 0., 1., 126., 127., 128., 129., 32766., 32768.,
                        32769., -1.9, -1.5, -1.1, -1., -0.9, -0.5, -0.1, 0.0, 0.1, 0.5, 0.9, 1.0, 1.1, 1.5, 1.9, 16056.8, 2.5,
                        1e30, -1e30, 3e9, -3e9, np.nan, np.inf, -np.inf, 126.5, 127.5, -126.5, -127.5], dtype=np.float32)
    x[:, :len(corners)] = corners
    _, Bt, bias, _, bq = util.make_qmm_case(2, 8, 256, 64)
    for aq in (1.0, -1.0, -0.49, 32.0):
        y, qa, acc = gpu_ctx.qmm_affine(x, Bt, bias, aq, bq, debug=True)
        yo, qao, acco = so.affine(x, Bt, bias, aq, bq, want=True)
        assert np.array_equal(qa, qao), aq
        assert np.array_equal(acc, acco) and np.array_equal(y, yo)


def test_saturation_cases_reported_separately(gpu_ctx):
    """Operands that overflow int16 pair sums on non-VNNI x86: the GPU accumulators equal the EXACT integer
    product (VNNI oracle); the maddubs-mode oracle differs and the number of saturating pairs is reported."""
    rng = np.random.RandomState(9)
    x = np.abs(rng.standard_normal((16, 256))).astype(np.float32) * 3
    Bt = rng.randint(60, 128, size=(64, 256)).astype(np.int8)
    bias = np.zeros(64, dtype=np.float32)
    aq, bq = 127.0 / 3.0, 127.0
    y, qa, acc = gpu_ctx.qmm_affine(x, Bt, bias, aq, bq, debug=True)
    n_sat = so.saturation_count(qa, Bt)
    assert n_sat > 0
    exact = (qa.astype(np.int64) + 127) @ Bt.astype(np.int64).T
    assert np.array_equal(acc.astype(np.int64), exact)
    assert np.array_equal(y, so.affine(x, Bt, bias, aq, bq, exact=True))
    assert not np.array_equal(acc, so.gemm_shifted(qa, Bt, exact=False))
    print(f"saturating adjacent-k pairs: {n_sat} of {16 * 64 * 128}")


def test_full_size_output_projection_exact_integer_product(gpu_ctx):
    """BASELINE-size GEMM (4096 x 256 x 32000): accumulators against an exact float64-BLAS integer product,
    f32 outputs against the oracle epilogue; plus linearity of the accumulator in A (size-independent)."""
    M, K, N = 4096, 256, 32000
    x, Bt, bias, aq, bq = util.make_qmm_case(77, M, K, N)
    y, qa, acc = gpu_ctx.qmm_affine(x, Bt, bias, aq, bq, debug=True)
    assert np.array_equal(qa, so.quantize(x, aq))
    exact = (qa.astype(np.float64) + 127.0) @ Bt.astype(np.float64).T
    assert np.array_equal(acc.astype(np.float64), exact)
    pb, _ = so.prepare_bias(Bt, bias, aq, bq)
    assert np.array_equal(y, so.unquantize(acc, pb, aq, bq))


def test_golden_qmm(gpu_ctx):
    g = np.load(os.path.join(util.GOLDEN, "qmm_cases.npz"))
    shapes = util.REFERENCE_GEMM_SHAPES + [(24, 256, 1536), (24, 1536, 256)]
    for i, (M, K, N) in enumerate(shapes):
        x, Bt, bias, aq, bq = util.make_qmm_case(100 + i, M, K, N)
        y, qa, _ = gpu_ctx.qmm_affine(x, Bt, bias, aq, bq, debug=True)
        assert np.array_equal(y, g[f"y_{i}"]), (M, K, N)
        u8 = (qa.astype(np.int16) + 127).astype(np.uint64)
        assert [int(u8.sum()), int((u8 * (np.arange(u8.size).reshape(u8.shape) % 251 + 1)).sum())] == g[f"qa_crc_{i}"].tolist()
    x, Bt, bias, aq, bq = util.make_qmm_case(200, 16, 256, 4096)
    idx = np.sort(np.random.RandomState(5).choice(4096, 512, replace=False)).astype(np.uint32)
    assert np.array_equal(gpu_ctx.qmm_affine(x, Bt, bias, aq, bq, idx), g["y_select"])


def test_shape_preconditions_are_errors(gpu_ctx):
    x, Bt, bias, aq, bq = util.make_qmm_case(1, 8, 256, 64)
    with pytest.raises(RuntimeError, match="precondition"):
        gpu_ctx.qmm_affine(x[:, :100], Bt[:, :100], bias, aq, bq)
    with pytest.raises(RuntimeError, match="precondition"):
        gpu_ctx.qmm_affine(x, Bt[:60], bias[:60], aq, bq)
    assert gpu_ctx.qmm_affine(x[:0], Bt, bias, aq, bq).shape == (0, 64)
