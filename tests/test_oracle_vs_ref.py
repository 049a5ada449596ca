"""CPU: pins the oracle port against the UNMODIFIED reference compiled in place (oracle/_ref/slimt_ref,
built by oracle/Makefile from /root/reference).  Everything here is BIT-EXACT."""
import numpy as np
import pytest

from oracle import slimt_oracle as so
from slimt_b200 import synth
import sb_testutil as util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/slimt_ref not built (needs /root/reference)")


@pytest.mark.parametrize("shape", util.REFERENCE_GEMM_SHAPES + util.HOTPATH_GEMM_SHAPES)
def test_qmm_affine_bit_exact(shape):
    M, K, N = shape
    x, Bt, bias, aq, bq = util.make_qmm_case(M + K + N, M, K, N)
    y_ref, qa_ref = util.ref_qmm(x, Bt, bias, aq, bq)
    y, qa, _ = so.affine(x, Bt, bias, aq, bq, want=True)
    assert np.array_equal(qa.astype(np.int16) + 127, qa_ref.astype(np.int16))  # PrepareA output (u8)
    assert np.array_equal(y, y_ref)


def test_qmm_affine_with_select_bit_exact():
    x, Bt, bias, aq, bq = util.make_qmm_case(11, 16, 256, 32000)
    idx = np.sort(np.random.RandomState(5).choice(32000, 2048, replace=False)).astype(np.uint32)
    y_ref, _ = util.ref_qmm(x, Bt, bias, aq, bq, indices=idx)
    assert np.array_equal(so.affine(x, Bt, bias, aq, bq, indices=idx), y_ref)


def test_qmm_corner_activations_bit_exact():
    x = np.zeros((8, 256), dtype=np.float32)
    vals = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 126.5, 127.5, -127.5, 1e30, -1e30, 3e9, -3e9, 200.0, -200.0, np.nan, 0.0], dtype=np.float32)
    x[:, :16] = vals
    _, Bt, bias, _, bq = util.make_qmm_case(2, 8, 256, 64)
    y_ref, qa_ref = util.ref_qmm(x, Bt, bias, 1.0, bq)
    y, qa, _ = so.affine(x, Bt, bias, 1.0, bq, want=True)
    assert np.array_equal(qa.astype(np.int16) + 127, qa_ref.astype(np.int16))
    assert np.array_equal(y, y_ref)


def test_saturating_isa_matches_maddubs_mode():
    """INTGEMM_CPUID=AVX512BW reproduces gemmology's maddubs arithmetic; the port's exact=False mode follows it."""
    rng = np.random.RandomState(9)
    x = np.abs(rng.standard_normal((8, 256))).astype(np.float32) * 3
    Bt = rng.randint(60, 128, size=(64, 256)).astype(np.int8)
    bias = np.zeros(64, dtype=np.float32)
    aq, bq = 127.0 / 3.0, 127.0
    qa = so.quantize(x, aq)
    assert so.saturation_count(qa, Bt) > 0
    try:
        y_bw, _ = util.ref_qmm(x, Bt, bias, aq, bq, env={"INTGEMM_CPUID": "AVX512BW"})
    except Exception:
        pytest.skip("host CPU cannot run the AVX512BW kernel")
    y_vnni, _ = util.ref_qmm(x, Bt, bias, aq, bq, env={"INTGEMM_CPUID": "AVX512VNNI"})
    assert np.array_equal(so.affine(x, Bt, bias, aq, bq, exact=True), y_vnni)
    assert np.array_equal(so.affine(x, Bt, bias, aq, bq, exact=False), y_bw)
    assert not np.array_equal(y_bw, y_vnni)


def test_f32_ops_bit_exact():
    rng = np.random.RandomState(1)
    x = (rng.standard_normal((64, 256)) * 2).astype(np.float32)
    s = (1 + 0.1 * rng.standard_normal(256)).astype(np.float32)
    b = (0.1 * rng.standard_normal(256)).astype(np.float32)
    assert np.array_equal(so.layer_norm(x, s, b).ravel(), util.ref_op("layer_norm", (64, 256), [x, s, b]))
    x = (rng.standard_normal((128, 40)) * 5).astype(np.float32)
    assert np.array_equal(so.softmax(x).ravel(), util.ref_op("softmax", (128, 40), [x]))
    xs = [(rng.standard_normal(4096) * 3).astype(np.float32) for _ in range(3)]
    assert np.array_equal(so.highway(*xs), util.ref_op("highway", (4096,), xs))
    assert np.array_equal(so.sinusoid(0, 64, 256).ravel(), util.ref_op("sinusoid", (0, 64, 256), []))


@pytest.mark.parametrize("dims", [(4, 8, 32, 32, 32), (4, 8, 1, 32, 32), (2, 8, 17, 17, 32), (2, 8, 40, 40, 64)])
def test_sdpa_bit_exact(dims):
    B, H, Tq, Tk, dh = dims
    E = H * dh
    rng = np.random.RandomState(sum(dims))
    q = (rng.standard_normal((B, Tq, E)) * 2).astype(np.float32)
    k = (rng.standard_normal((B, Tk, E)) * 2).astype(np.float32)
    v = rng.standard_normal((B, Tk, E)).astype(np.float32)
    lens = rng.randint(1, Tk + 1, size=B)
    mask = ((np.arange(Tk)[None] >= lens[:, None]) * np.float32(-99999999.0)).astype(np.float32)
    out, attn = so.sdpa(q, k, v, mask, H)
    r = util.ref_op("sdpa", dims, [q, k, v, mask])
    assert np.array_equal(out.ravel(), r[:out.size]) and np.array_equal(attn.ravel(), r[out.size:])


@pytest.mark.parametrize("use_shortlist", [False, True])
def test_forward_bit_exact(tiny_model, shortlist_assets, use_shortlist):
    path, items = tiny_model
    sents = synth.make_sentences(6, (2, 12), seed=21)
    tokens, lengths = util.pad_batch(sents)
    sl = None
    if use_shortlist:
        fr, offs, lists = shortlist_assets[1]
        sl = so.shortlist_generate(np.concatenate(sents), fr, offs, lists, 32000)
    ref = util.ref_forward(path, sents, shortlist=sl, dump=True)
    out = so.Oracle(items).forward(tokens, lengths, shortlist=sl, keep=True)
    assert np.array_equal(out["trace"]["embed"] if "trace" in out else so.Oracle(items).embed(tokens), ref["embed"])
    assert np.array_equal(out["encoder_out"], ref["encoder_out"])
    assert np.array_equal(out["step_tokens"], ref["step_tokens"])
    assert out["sentences"] == ref["sentences"]
    for s in range(len(ref["step_tokens"])):
        assert np.array_equal(out["logits"][s], ref["logits"][s]), s
        assert np.array_equal(out["attn"][s], ref["attn"][s]), s


def test_forward_long_sentences_bit_exact(tiny_model):
    """Sentences longer than 32 tokens (positions, softmax rows and decode lengths beyond the short cases above):
    the restatement still equals the reference bit for bit."""
    path, items = tiny_model
    sents = synth.make_sentences(3, (20, 44), seed=57)
    sents[0] = synth.make_sentences(1, 44, seed=58)[0]
    tokens, lengths = util.pad_batch(sents)
    ref = util.ref_forward(path, sents, dump=True)
    out = so.Oracle(items).forward(tokens, lengths, keep=True)
    assert np.array_equal(out["encoder_out"], ref["encoder_out"])
    assert np.array_equal(out["step_tokens"], ref["step_tokens"])
    assert out["sentences"] == ref["sentences"]
    for s in range(len(ref["step_tokens"])):
        assert np.array_equal(out["logits"][s], ref["logits"][s]), s
        assert np.array_equal(out["attn"][s], ref["attn"][s]), s


def test_forward_base_shaped_bit_exact(tmp_path):
    """BASELINE configs[2] shapes (emb 512, ffn 2048, head size 64): restatement == reference."""
    path = str(tmp_path / "base.bin")
    dims = synth.ModelDims(emb=512, ffn=2048, vocab=8000)
    synth.write_model(path, synth.make_params(dims, seed=5))
    items = synth.read_model(path)
    sents = synth.make_sentences(4, (2, 10), vocab=8000, seed=6)
    tokens, lengths = util.pad_batch(sents)
    ref = util.ref_forward(path, sents, dump=True)
    out = so.Oracle(items).forward(tokens, lengths, keep=True)
    assert np.array_equal(out["encoder_out"], ref["encoder_out"])
    assert np.array_equal(out["step_tokens"], ref["step_tokens"])
    for s in range(len(ref["step_tokens"])):
        assert np.array_equal(out["logits"][s], ref["logits"][s]), s
        assert np.array_equal(out["attn"][s], ref["attn"][s]), s


def test_forward_with_eos_and_teacher_forcing(eos_model):
    path, items = eos_model
    sents = synth.make_sentences(8, (4, 10), seed=33)
    tokens, lengths = util.pad_batch(sents)
    ref = util.ref_forward(path, sents)
    out = so.Oracle(items).forward(tokens, lengths)
    assert out["sentences"] == ref["sentences"]
    assert np.array_equal(out["step_tokens"], ref["step_tokens"])
    assert len({len(s) for s in ref["sentences"]}) > 1, "fixture should make sentences finish at different steps"
    T = tokens.shape[1]
    forced = np.random.RandomState(4).randint(1, 32000, size=(int(1.5 * T), 8)).astype(np.uint32)
    ref_f = util.ref_forward(path, sents, forced=forced, dump=True)
    out_f = so.Oracle(items).forward(tokens, lengths, forced=forced, keep=True)
    n = len(ref_f["step_tokens"])
    assert np.array_equal(out_f["step_tokens"][:n], ref_f["step_tokens"])
    assert all(np.array_equal(out_f["logits"][s], ref_f["logits"][s]) for s in range(n))
