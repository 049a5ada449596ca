"""GPU: BASELINE.json's own sizes against token goldens produced by the UNMODIFIED reference (tests/golden/
make_golden_large.py ran oracle/_ref in the build container; the files travel, /root/reference does not), plus the
greedy tie rule of the fused output GEMM.  Everything goes through the production path: the fused output-GEMM + argmax
kernel, the shortlist gather, and for the mixed request slimt_b200_translate."""
import importlib.util
import os
import struct

import numpy as np
import pytest

import sb_testutil as util
from oracle import slimt_oracle as so
from slimt_b200 import capi, synth

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("make_golden_large", os.path.join(util.GOLDEN, "make_golden_large.py"))
_mgl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mgl)


def _load(name, tmp_path_factory):
    g = np.load(os.path.join(util.GOLDEN, f"large_{name}.npz"))
    tmp = tmp_path_factory.mktemp("large")
    path, sha, sents, sl = _mgl.case_assets(name, tmp)
    assert sha == g["model_sha256"].tobytes(), "synthetic model generator drifted from the golden fixture"
    return g, path, sents, sl


@pytest.mark.parametrize("name", ["tiny_shortlist_4096x32", "tiny_full_4096x32", "base_shortlist_1024x32"])
def test_baseline_size_tokens_equal_reference(gpu_ctx, name, tmp_path_factory):
    """configs[1] / [3] / [2] in ONE batch, exactly as the reference ran them: every token of every step equal."""
    g, path, sents, sl = _load(name, tmp_path_factory)
    dims = getattr(synth, _mgl.LARGE_CASES[name]["dims"])
    m = capi.Model(gpu_ctx, open(path, "rb").read())
    tokens, lengths = util.pad_batch(sents)
    words = so.shortlist_generate(np.concatenate(sents), *sl, dims.vocab) if sl is not None else None
    assert int(g["shortlist_size"][0]) == (0 if words is None else len(words))
    out = m.forward(tokens, lengths, limit_factor=_mgl.LIMIT, shortlist=words)
    want = g["step_tokens"].astype(np.uint32)
    assert out["steps"] == want.shape[0]
    assert np.array_equal(out["step_tokens"], want), f"{(out['step_tokens'] != want).sum()} of {want.size} tokens differ"
    assert out["target_tokens"] == int(g["sentence_lengths"].sum())
    m.close()


def test_mixed_length_translate_equals_reference(gpu_ctx, tmp_path_factory):
    """configs[4] in miniature: one translate request with lengths 8-256 and max_words 2^17; the library's Batcher must
    form the batches the reference saw and every sentence must come back as the reference decoded it."""
    name = "mixed_translate"
    g, path, sents, sl = _load(name, tmp_path_factory)
    c = _mgl.LARGE_CASES[name]
    sl_path = os.path.join(os.path.dirname(path), "lex.s2t.bin")
    synth.write_shortlist(sl_path, *sl, best=100)
    m = capi.Model(gpu_ctx, open(path, "rb").read())
    outs, stats = m.translate(sents, max_words=c["max_words"], limit_factor=_mgl.LIMIT, shortlist_bin=open(sl_path, "rb").read())
    assert stats["batches"] == len(g["batch_widths"])
    offs = g["offsets"].astype(np.int64)
    bad = [i for i in range(len(sents)) if outs[i].tolist() != g["tokens"][offs[i]:offs[i + 1]].astype(np.uint32).tolist()]
    assert not bad, f"{len(bad)} of {len(sents)} sentences differ from the reference (first: {bad[:5]})"
    assert stats["target_tokens"] == int(offs[-1])
    m.close()


def _periodic_model(path, period):
    """tiny11 whose output layer repeats with `period`: row n of Wemb and its bias equal row n % period, so every logit
    value occurs in V / period columns, spread over different 256-column tiles and different CTAs of the output GEMM."""
    items = synth.make_params(synth.TINY, seed=1234)
    typ, shape, blob = items["Wemb"]
    V, E = shape
    q = np.frombuffer(blob[:V * E], dtype=np.int8).reshape(V, E).copy()
    q[:] = q[np.arange(V) % period]
    items["Wemb"] = (typ, shape, q.tobytes() + blob[V * E:])
    typ, shape, blob = items["decoder_ff_logit_out_b"]
    b = np.frombuffer(blob, dtype=np.float32).copy()
    b[:] = b[np.arange(V) % period]
    b[0] = b[period]  # the synthetic EOS bias would break the tie at column 0
    items["decoder_ff_logit_out_b"] = (typ, shape, b.tobytes())
    synth.write_model(path, items)


@pytest.mark.parametrize("period,B", [(256, 40), (1000, 200)])
def test_fused_argmax_ties_take_the_lowest_index(gpu_ctx, tmp_path, period, B):
    """greedy_sample keeps the FIRST strict maximum (Transformer.cc:291-297, 323-329).  With a periodic output layer
    every row's maximum is attained in V / period columns that live in different column tiles (period 256: the same
    position of every tile; period 1000: different positions and chunks), reduced by different CTAs in no fixed order."""
    path = str(tmp_path / f"periodic_{period}.bin")
    _periodic_model(path, period)
    m = capi.Model(gpu_ctx, open(path, "rb").read())
    orc = so.Oracle(synth.read_model(path))
    sents = synth.make_sentences(B, (2, 8), seed=period)
    tokens, lengths = util.pad_batch(sents)
    ref = orc.forward(tokens, lengths, keep=True)
    lg = ref["logits"][0]
    top = np.sort(lg, axis=-1)[:, -2:]
    assert (top[:, 0] == top[:, 1]).all(), "the construction must produce tied maxima"
    assert (np.asarray(ref["step_tokens"]) < period).all()
    fused = m.forward(tokens, lengths)
    assert np.array_equal(fused["step_tokens"], ref["step_tokens"])
    # with a shortlist the candidates are the sorted ids: the lowest position is still the lowest id
    fr, offs, lists = synth.make_shortlist(vocab=32000, frequent=100, best=100, seed=7)
    sl = so.shortlist_generate(np.concatenate(sents), fr, offs, lists, 32000)
    ref_sl = orc.forward(tokens, lengths, shortlist=sl)
    assert np.array_equal(m.forward(tokens, lengths, shortlist=sl)["step_tokens"], ref_sl["step_tokens"])
    m.close()


def test_replicas_and_lanes_do_not_change_the_output(gpu_ctx, tiny_model, shortlist_assets, monkeypatch):
    """slimt_b200_translate_multi: one Batcher, batches dealt to whichever lane is free.  Two replicas (here on the same
    GPU; on a multi-GPU box one per device) with two lanes each must return exactly what one replica on one lane does,
    alignments included."""
    path, _ = tiny_model
    sl_bin = open(shortlist_assets[0], "rb").read()
    blob = open(path, "rb").read()
    a, b = capi.Model(gpu_ctx, blob), capi.Model(gpu_ctx, blob)
    sents = synth.make_sentences(300, (2, 40), seed=123)
    monkeypatch.setenv("SLIMT_B200_LANES", "1")
    one, st1 = a.translate(sents, max_words=512, shortlist_bin=sl_bin, want_alignments=True)
    monkeypatch.setenv("SLIMT_B200_LANES", "2")
    two, st2 = a.translate(sents, max_words=512, shortlist_bin=sl_bin, want_alignments=True, replicas=[a, b])
    assert st1["batches"] == st2["batches"] > 8
    assert all(np.array_equal(x, y) for x, y in zip(one, two))
    assert all(np.array_equal(x, y) for x, y in zip(st1["alignments"], st2["alignments"]))
    assert st2["target_tokens"] == st1["target_tokens"] and st2["kernel_launches"] == st1["kernel_launches"]
    a.close(), b.close()


def test_replicas_on_two_gpus_equal_one(gpu_ctx, tiny_model, shortlist_assets):
    """The multi-GPU shape of the service: one replica per device, batches dealt by one Batcher in one process.  Needs a
    second GPU (skipped on a single-GPU box; exercised by tools/gpu_multi_r2.sh on the multi-GPU boxes)."""
    try:
        ctx1 = capi.Context(1)
    except RuntimeError:
        pytest.skip("needs two GPUs")
    path, _ = tiny_model
    sl_bin = open(shortlist_assets[0], "rb").read()
    blob = open(path, "rb").read()
    a, b = capi.Model(gpu_ctx, blob), capi.Model(ctx1, blob)
    sents = synth.make_sentences(400, (2, 70), seed=321)
    one, st1 = a.translate(sents, max_words=1024, shortlist_bin=sl_bin)
    two, st2 = a.translate(sents, max_words=1024, shortlist_bin=sl_bin, replicas=[a, b])
    assert st1["batches"] == st2["batches"] > 8
    assert all(np.array_equal(x, y) for x, y in zip(one, two))
    assert st2["target_tokens"] == st1["target_tokens"]
    a.close(), b.close(), ctx1.close()


def test_service_alignments_equal_oracle(gpu_ctx, tiny_model, shortlist_assets):
    """Response.alignments through the service call (Model.cc:84-108): per target token, head 0 of the last decoder
    layer's cross-attention over the sentence's own source tokens, for the tokens record() kept."""
    path, items = tiny_model
    sl_path, (fr, offs, lists) = shortlist_assets
    m = capi.Model(gpu_ctx, open(path, "rb").read())
    orc = so.Oracle(items)
    sents = synth.make_sentences(24, (2, 12), seed=77)
    outs, st = m.translate(sents, max_words=96, shortlist_bin=open(sl_path, "rb").read(), want_alignments=True)
    for ids, width in util.batcher_generate_py([len(s) for s in sents], 96):
        chunk = [sents[i] for i in ids]
        tk, ln = util.pad_batch(chunk)
        sl = so.shortlist_generate(np.concatenate(chunk), fr, offs, lists, 32000)
        ref = orc.forward(tk, ln, shortlist=sl, keep=True)
        for r, i in enumerate(ids):
            assert outs[i].tolist() == ref["sentences"][r]
            want = np.stack([ref["attn"][s][r, 0, 0, :ln[r]] for s in range(len(ref["sentences"][r]))])
            assert np.array_equal(st["alignments"][i], want), f"sentence {i}"
    m.close()


def test_eos_and_pad_ids_come_from_the_model_config(gpu_ctx, tmp_path):
    """Vocabulary::eos_id() / pad_id() are configuration, not constants: a model whose EOS is id 7 stops on 7."""
    path = str(tmp_path / "eos7.bin")
    items = synth.make_params(synth.TINY, seed=4321, eos_bias=0.0)
    typ, shape, blob = items["decoder_ff_logit_out_b"]
    b = np.frombuffer(blob, dtype=np.float32).copy()
    b[7] = 4.8
    items["decoder_ff_logit_out_b"] = (typ, shape, b.tobytes())
    synth.write_model(path, items)
    blob = open(path, "rb").read()
    sents = synth.make_sentences(8, (4, 10), seed=33)
    tokens, lengths = util.pad_batch(sents)
    m0, m7 = capi.Model(gpu_ctx, blob), capi.Model(gpu_ctx, blob, eos_id=7, pad_id=7)
    a, b7 = m0.forward(tokens, lengths), m7.forward(tokens, lengths)
    n = b7["steps"]
    assert np.array_equal(a["step_tokens"][:n], b7["step_tokens"])  # same arithmetic, different bookkeeping
    steps = b7["step_tokens"]
    expect = sum(int(np.argmax(steps[:, r] == 7)) + 1 if (steps[:, r] == 7).any() else n for r in range(len(sents)))
    assert b7["target_tokens"] == expect
    assert (steps == 7).any(), "the construction must emit the configured EOS"
    outs, _ = m7.translate(sents, max_words=4096)
    assert all(o[-1] == 7 or len(o) == max(1, int(np.float32(1.5) * np.float32(tokens.shape[1]))) for o in outs)
    m0.close(), m7.close()


def test_at_least_one_step_even_when_the_limit_rounds_to_zero(gpu_ctx, tiny_model):
    """Model::decode runs its first step before the `i < limit_factor * T` loop (Model.cc:145-161)."""
    m = capi.Model(gpu_ctx, open(tiny_model[0], "rb").read())
    out = m.forward(np.array([[5], [9]], np.uint32), np.array([1, 1], np.uint32), limit_factor=0.5)
    assert out["steps"] == 1 and out["step_tokens"].shape == (1, 2) and out["target_tokens"] == 2
    m.close()
