"""Replays the golden cases of tests/golden/make_golden.py against any implementation of forward()."""
import hashlib
import importlib.util
import os

import numpy as np

import sb_testutil as util
from oracle import slimt_oracle as so
from slimt_b200 import synth

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(util.GOLDEN, "make_golden.py"))
_mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mg)
FORWARD_CASES = _mg.FORWARD_CASES


def load_case(name, tmpdir):
    """Returns (model_path, tokens, lengths, shortlist, forced, golden npz)."""
    pk, sk, use_sl, forced_seed = FORWARD_CASES[name]
    g = np.load(os.path.join(util.GOLDEN, f"forward_{name}.npz"))
    path = os.path.join(str(tmpdir), f"golden_{name}.bin")
    synth.write_model(path, synth.make_params(synth.TINY, **pk))
    sha = hashlib.sha256(open(path, "rb").read()).digest()
    assert sha == g["model_sha256"].tobytes(), "synthetic model generator drifted from the golden fixture"
    sents = synth.make_sentences(**sk)
    tokens, lengths = util.pad_batch(sents)
    sl = None
    if use_sl:
        fr, offs, lists = synth.make_shortlist(vocab=32000, frequent=100, best=100, seed=7)
        sl = so.shortlist_generate(np.concatenate(sents), fr, offs, lists, 32000)
    forced = g["forced"] if "forced" in g.files else None
    return path, tokens, lengths, sl, forced, g


def check_against_golden(g, lengths, step_tokens, encoder_out, logits, align=None):
    """Bit-exact comparison of an implementation's outputs with reference-generated vectors."""
    steps = len(g["step_tokens"])
    assert len(step_tokens) == steps
    assert np.array_equal(np.asarray(step_tokens), g["step_tokens"])
    T = encoder_out.shape[1]
    valid = np.arange(T)[None, :] < np.asarray(lengths)[:, None]
    assert np.array_equal(encoder_out[valid], g["encoder_out"][valid])
    logits = np.asarray(logits)
    assert np.array_equal(logits[..., ::97], g["logits_strided"])
    assert np.array_equal(np.take_along_axis(logits, g["logits_topk_idx"].astype(np.int64), axis=-1), g["logits_topk_val"])
    assert np.array_equal(logits.astype(np.float64).sum(axis=-1), g["logits_sum64"])
    if align is not None:
        assert np.array_equal(np.asarray(align)[:, valid], g["align_head0"][:, valid])
