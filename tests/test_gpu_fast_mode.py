"""GPU: the tolerance arithmetic mode (slimt_b200_ctx_set_math(ctx, 1) / SLIMT_B200_MATH=fast).

What this mode is and is not.  Every kernel keeps f32 / int32 arithmetic; the mode contracts the dequantisation into
FMAs, reduces LayerNorm sums as shuffle trees, takes softmax / sigmoid through ex2 / rcp and makes the greedy choice on
an integer proxy of the logit.  Each of those perturbs a float by a few ulp.  Between every pair of GEMMs the path
re-quantises to int8 with a step of 1/21 of an activation's standard deviation, so a few-ulp perturbation is invisible
unless it carries a value across a rounding boundary -- then one int8 operand changes by one step, and the random-init
synthetic model (weights scaled so that outputs depend on the input) amplifies that through its
layers to a few percent of the logit scale.  Measured here: most logit rows come out BIT-IDENTICAL to the exact mode,
the rest differ by ~1e-2 of their scale; once a token differs the sentence's later tokens differ too.  BASELINE.json's
bars (logits rtol 1e-3, >= 99 % tokens) are therefore met by the EXACT mode only (every other GPU test: 100 %), which
stays the default and the headline; the reference's own two code paths (AVX512-VNNI vs AVX512BW) agree on 96 % of the
tokens of this model (bench.py cpu_baseline.saturation).  This file pins what the tolerance mode does deliver, so that a
regression (a real arithmetic error rather than rounding noise) is caught."""
import os

import numpy as np
import pytest

import golden_cases
import sb_testutil as util
from oracle import slimt_oracle as so
from slimt_b200 import capi, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-3


@pytest.fixture(scope="module")
def fast_ctx():
    ctx = capi.Context(0)
    ctx.set_math(True)
    assert ctx.math() == "fast"
    yield ctx
    ctx.close()


def _row_err(fast, ref):
    fast, ref = np.asarray(fast, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max(axis=-1)
    return np.abs(fast - ref).max(axis=-1) / scale


def test_teacher_forced_logits_within_rtol(fast_ctx, tiny_model):
    """Same inputs at every step (teacher forcing), so the comparison is of arithmetic, not of diverging histories."""
    path, items = tiny_model
    m = capi.Model(fast_ctx, open(path, "rb").read())
    orc = so.Oracle(items)
    sents = synth.make_sentences(24, (3, 30), seed=5)
    tokens, lengths = util.pad_batch(sents)
    steps = int(np.float32(1.5) * np.float32(tokens.shape[1]))
    forced = np.random.RandomState(3).randint(1, 32000, size=(steps, len(sents))).astype(np.uint32)
    ref = orc.forward(tokens, lengths, forced=forced, keep=True)
    out = m.forward(tokens, lengths, forced=forced, want_logits=True, want_encoder=True, want_alignment=True)
    err = np.stack([_row_err(out["logits"][s], ref["logits"][s]) for s in range(out["steps"])])
    print(f"fast-mode logits: row error / row scale  median {np.median(err):.2e}  p90 {np.quantile(err, 0.9):.2e}  "
          f"max {err.max():.2e};  rows within 1e-3: {(err <= RTOL).mean():.3f}, bit-identical rows: {(err == 0).mean():.3f}")
    # rounding noise only: rows untouched by an int8 flip are (nearly) exact, flipped rows stay within a few percent
    assert np.median(err) <= RTOL and err.max() <= 0.2
    T = tokens.shape[1]
    valid = np.arange(T)[None, :] < lengths[:, None]
    enc_scale = np.abs(ref["encoder_out"][valid]).max()
    enc_err = np.abs(out["encoder_out"][valid] - ref["encoder_out"][valid]) / enc_scale
    assert np.median(enc_err) <= 1e-5 and (enc_err <= RTOL).mean() >= 0.5
    a = np.stack([x[:, 0, 0, :] for x in ref["attn"]])
    assert np.median(np.abs(out["alignment"][:, valid] - a[:, valid])) <= RTOL  # probabilities: absolute
    # the fused integer argmax must agree with the argmax of the same mode's float logits except at near-ties
    fused = m.forward(tokens, lengths, forced=forced)
    agree = (fused["step_tokens"] == out["step_tokens"]).mean()
    assert agree >= 0.999, agree
    m.close()


@pytest.mark.parametrize("name", sorted(golden_cases.FORWARD_CASES))
def test_reference_goldens_within_tolerance(fast_ctx, name, tmp_path):
    """The small reference-generated cases (plain, shortlist, EOS bookkeeping, teacher forcing, sentences of 33-50 tokens)."""
    path, tokens, lengths, sl, forced, g = golden_cases.load_case(name, tmp_path)
    m = capi.Model(fast_ctx, open(path, "rb").read())
    out = m.forward(tokens, lengths, shortlist=sl, forced=forced, want_logits=True)
    want = np.asarray(g["step_tokens"])
    n = min(out["steps"], len(want))
    errs = []
    same_so_far = np.ones(want.shape[1], dtype=bool)
    for s_ in range(n):
        top = np.take_along_axis(np.asarray(out["logits"][s_]), g["logits_topk_idx"][s_].astype(np.int64), axis=-1)
        scale = np.abs(g["logits_topk_val"][s_]).max(axis=-1)
        e = np.abs(top - g["logits_topk_val"][s_]).max(axis=-1) / scale
        # teacher forcing keeps every row comparable; free running only until a sentence's history first differs
        errs.extend(e.tolist() if forced is not None else e[same_so_far].tolist())
        same_so_far &= out["step_tokens"][s_] == want[s_]
    errs = np.asarray(errs)
    print(f"{name}: rows within 1e-3: {(errs <= RTOL).mean():.3f}, max {errs.max():.2e}; tokens equal {(out['step_tokens'][:n] == want[:n]).mean():.3f}")
    # long sentences have more quantisation points per sentence: every sentence of the `long` case carries a flip
    assert errs.max() <= 0.2
    assert (out["step_tokens"][:n] == want[:n]).mean() >= 0.5
    m.close()


@pytest.mark.parametrize("name", ["tiny_shortlist_4096x32", "tiny_full_4096x32", "base_shortlist_1024x32"])
def test_baseline_size_token_agreement_with_reference(fast_ctx, name, tmp_path_factory):
    """Token agreement with the reference at BASELINE.json's own batch sizes, through the production path: reported, and
    guarded against collapse (an arithmetic error would drop it to the chance level of ~0)."""
    import test_gpu_large_golden as big
    g, path, sents, sl = big._load(name, tmp_path_factory)
    dims = getattr(synth, big._mgl.LARGE_CASES[name]["dims"])
    m = capi.Model(fast_ctx, open(path, "rb").read())
    tokens, lengths = util.pad_batch(sents)
    words = so.shortlist_generate(np.concatenate(sents), *sl, dims.vocab) if sl is not None else None
    out = m.forward(tokens, lengths, limit_factor=big._mgl.LIMIT, shortlist=words)
    want = g["step_tokens"].astype(np.uint32)
    assert out["steps"] == want.shape[0]
    tok = (out["step_tokens"] == want).mean()
    sent = (out["step_tokens"] == want).all(axis=0).mean()
    print(f"{name}: fast mode vs reference: {tok:.5f} of step tokens, {sent:.5f} of sentences identical")
    assert tok >= 0.75
    m.close()


def test_mixed_translate_token_agreement(fast_ctx, tmp_path_factory):
    """The mixed-length request (lengths 8-256: the split self-attention and cached cross-attention kernels stay exact,
    the row kernels and the output layer run in tolerance mode)."""
    import test_gpu_large_golden as big
    name = "mixed_translate"
    g, path, sents, sl = big._load(name, tmp_path_factory)
    c = big._mgl.LARGE_CASES[name]
    sl_path = os.path.join(os.path.dirname(path), "lex.s2t.bin")
    synth.write_shortlist(sl_path, *sl, best=100)
    m = capi.Model(fast_ctx, open(path, "rb").read())
    outs, stats = m.translate(sents, max_words=c["max_words"], limit_factor=big._mgl.LIMIT, shortlist_bin=open(sl_path, "rb").read())
    offs = g["offsets"].astype(np.int64)
    same = total = 0
    for i in range(len(sents)):
        want = g["tokens"][offs[i]:offs[i + 1]].astype(np.uint32)
        n = min(len(want), len(outs[i]))
        same += int((outs[i][:n] == want[:n]).sum())
        total += max(len(want), len(outs[i]))
    print(f"mixed translate: fast mode vs reference: {same / total:.5f} of tokens")
    assert same / total >= 0.6
    m.close()


def test_mode_is_per_context_and_switchable(gpu_ctx, tiny_model):
    """set_math flips one context; switching back restores bit-exact results."""
    path, items = tiny_model
    m = capi.Model(gpu_ctx, open(path, "rb").read())
    sents = synth.make_sentences(8, (3, 9), seed=2)
    tokens, lengths = util.pad_batch(sents)
    exact = m.forward(tokens, lengths, want_logits=True)
    gpu_ctx.set_math(True)
    fast = m.forward(tokens, lengths, want_logits=True)
    gpu_ctx.set_math(False)
    again = m.forward(tokens, lengths, want_logits=True)
    assert np.array_equal(exact["logits"], again["logits"])
    assert not np.array_equal(exact["logits"], fast["logits"])
    assert np.median(_row_err(fast["logits"][0], exact["logits"][0])) <= RTOL
    m.close()
