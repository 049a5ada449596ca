"""Regenerates the committed golden vectors by running the UNMODIFIED reference (oracle/_ref/slimt_ref,
compiled in place from /root/reference) in THIS container.  The GPU box has no /root/reference, so the
-m gpu tests and the CPU oracle tests compare against these files.

    python tests/golden/make_golden.py

Model weights are not stored: they are regenerated from the seed (numpy's frozen legacy RandomState); the
sha256 of the model image is stored so that generator drift is detected instead of silently mis-compared.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sb_testutil as util  # noqa: E402
from oracle import slimt_oracle as so  # noqa: E402
from slimt_b200 import synth  # noqa: E402


def topk(logits, k=8):
    idx = np.argsort(-logits, axis=-1, kind="stable")[..., :k]
    return idx.astype(np.uint32), np.take_along_axis(logits, idx, axis=-1)


def forward_case(name, params_kw, sent_kw, use_shortlist, forced_seed=None):
    path = f"/tmp/golden_{name}.bin"
    synth.write_model(path, synth.make_params(synth.TINY, **params_kw))
    sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
    sents = synth.make_sentences(**sent_kw)
    sl = None
    if use_shortlist:
        fr, offs, lists = synth.make_shortlist(vocab=32000, frequent=100, best=100, seed=7)
        sl = so.shortlist_generate(np.concatenate(sents), fr, offs, lists, 32000)
    T = max(len(s) for s in sents)
    forced = None
    if forced_seed is not None:
        forced = np.random.RandomState(forced_seed).randint(1, 32000, size=(int(np.float32(1.5) * np.float32(T)), len(sents))).astype(np.uint32)
    ref = util.ref_forward(path, sents, shortlist=sl, forced=forced, dump=True)
    logits = np.stack(ref["logits"])  # [steps, B, N]
    ti, tv = topk(logits)
    out = {
        "model_sha256": np.frombuffer(bytes.fromhex(sha), dtype=np.uint8),
        "encoder_out": ref["encoder_out"],
        "step_tokens": ref["step_tokens"],
        "sentence_lengths": np.array([len(s) for s in ref["sentences"]], dtype=np.uint32),
        "logits_topk_idx": ti, "logits_topk_val": tv,
        "logits_strided": logits[..., ::97].copy(),
        "logits_sum64": logits.astype(np.float64).sum(axis=-1),
        "align_head0": np.stack([a[:, 0, 0, :] for a in ref["attn"]]),
    }
    if forced is not None:
        out["forced"] = forced
    np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **out)
    print(name, "steps", len(ref["step_tokens"]), "lens", out["sentence_lengths"].tolist())


def qmm_cases():
    out = {}
    for i, (M, K, N) in enumerate(util.REFERENCE_GEMM_SHAPES + [(24, 256, 1536), (24, 1536, 256)]):
        x, Bt, bias, aq, bq = util.make_qmm_case(100 + i, M, K, N)
        y, qa = util.ref_qmm(x, Bt, bias, aq, bq)
        out[f"y_{i}"] = y
        out[f"qa_crc_{i}"] = np.array([int(qa.astype(np.uint64).sum()), int((qa.astype(np.uint64) * (np.arange(qa.size).reshape(qa.shape) % 251 + 1)).sum())], dtype=np.uint64)
    x, Bt, bias, aq, bq = util.make_qmm_case(200, 16, 256, 4096)
    idx = np.sort(np.random.RandomState(5).choice(4096, 512, replace=False)).astype(np.uint32)
    out["y_select"], _ = util.ref_qmm(x, Bt, bias, aq, bq, indices=idx)
    np.savez_compressed(os.path.join(HERE, "qmm_cases.npz"), **out)
    print("qmm cases", len(out))


# Definitions shared with the tests that replay these cases.
FORWARD_CASES = {
    "plain": (dict(seed=1234), dict(n=4, length=(3, 9), seed=21), False, None),
    "shortlist": (dict(seed=1234), dict(n=5, length=(2, 8), seed=22), True, None),
    "eos": (dict(seed=4321, eos_bias=4.8), dict(n=8, length=(4, 10), seed=33), False, None),
    "forced": (dict(seed=1234), dict(n=3, length=(5, 7), seed=23), False, 4),
    # sentences of 33-50 tokens: positions past 32, the two-key-block recompute cross-attention, T = 64 encoder attention
    "long": (dict(seed=1234), dict(n=3, length=(33, 50), seed=24), False, None),
}

if __name__ == "__main__":
    assert util.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    only = sys.argv[1:]  # optional: regenerate just the named forward cases
    for name, (pk, sk, sl, fs) in FORWARD_CASES.items():
        if not only or name in only:
            forward_case(name, pk, sk, sl, fs)
    if not only:
        qmm_cases()
