"""Text in, text out on the GPU through the C++ services (include/slimt_b200.hh + slimt_b200_text.hh): Model from a
Package with a sentencepiece vocabulary, Blocking::translate / pivot on strings, Async on a string.

The checker composes two pins: the text front half is byte-identical to the reference's own (tests/test_text_front.py),
the translation of word ids is token- and alignment-identical to the oracle (the other GPU suites, through the same
C-ABI call this test uses via ctypes).  Here: the segments the service batches equal the sentencepiece wheel's ids, the
target words equal the ctypes path's on those segments, the target text and its annotation equal the wheel's decode laid
out as Request::complete does (Request.cc:133-169), and the alignments are the C-ABI's rows."""
import os
import subprocess

import numpy as np
import pytest

from slimt_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_text_test")
SPM = os.path.join(ROOT, "tests", "golden", "text", "spm_unigram.model")

TEXTS = [
    "Hello world. This is one line.\nSecond line here!\n\nFourth after an empty line",
    "  ça va, Ñandú?  \n北京大学 Beijing 2024年 😀\n",
    "a considerably longer line that has to be wrapped several times because the limit is small for this test\nshort",
]


def _hx(s):
    return s.encode("utf-8").hex() if s else "-"


def _floats(field):
    out = []
    for group in field.split("[")[1:]:
        out.append(np.array([float.fromhex(x) for x in group.split("]")[0].split()], dtype=np.float32))
    return out


def test_text_service_builds_and_refuses_without_gpu(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "slimt_b200", "csrc"), "host_text"])
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    except ImportError:
        pass
    path = str(tmp_path / "m.bin")
    synth.write_model(path, synth.make_params(synth.ModelDims(vocab=1024), seed=77))
    r = subprocess.run([BIN, path, SPM, "-", "sentence", "16", "4096", _hx("Hello world.")], capture_output=True, text=True)
    assert r.returncode == 1 and ("no CPU fallback" in r.stderr or "CUDA" in r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("wrap_length,use_shortlist", [(16, True), (128, False)])
def test_text_in_text_out(gpu_ctx, tmp_path, wrap_length, use_shortlist):
    spm = pytest.importorskip("sentencepiece")
    from slimt_b200 import capi
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "slimt_b200", "csrc"), "host_text"])
    sp = spm.SentencePieceProcessor(model_file=SPM)
    V = sp.get_piece_size()
    assert V == 1024 and sp.eos_id() == 0
    path = str(tmp_path / "m.bin")
    synth.write_model(path, synth.make_params(synth.ModelDims(vocab=V), seed=77, eos_bias=3.0))
    sl_path, sl_bin = "-", None
    if use_shortlist:
        fr, offs, lists = synth.make_shortlist(vocab=V, frequent=40, best=24, seed=3)
        sl_path = str(tmp_path / "lex.bin")
        synth.write_shortlist(sl_path, fr, offs, lists, best=24)
        sl_bin = open(sl_path, "rb").read()
    max_words = 4096
    r = subprocess.run([BIN, path, SPM, sl_path, "sentence", str(wrap_length), str(max_words)] + [_hx(t) for t in TEXTS],
                       capture_output=True, text=True, check=True)
    lines = r.stdout.strip().split("\n")
    assert lines[0] == f"vocab size={V} eos=0 pad=0"
    blocking = [l for l in lines if l.startswith("blocking ")]
    segs = [l for l in lines if l.startswith("seg")]
    assert len(blocking) == 3 * len(TEXTS) and len(segs) == len(TEXTS)

    # 1. segments = lines of the text ("sentence" mode), tokenised by sentencepiece, cut at wrap_length - 1 words + EOS
    want_segments = []
    for text in TEXTS:
        mine = []
        for line in text.split("\n"):
            ids = sp.encode(line, out_type=int)
            for o in range(0, len(ids), wrap_length - 1):
                mine.append(ids[o:o + wrap_length - 1] + [0])
        want_segments.append(mine)
    got_segments = [[list(map(int, g.split(","))) for g in l.split()[1:]] for l in segs]
    assert got_segments == want_segments

    # 2. the words: the same request through the C ABI with ctypes (one pool of segments, the service's max_words)
    model = capi.Model(gpu_ctx, open(path, "rb").read())
    pool = [np.asarray(s, dtype=np.uint32) for segs_ in want_segments for s in segs_]
    outs, st = model.translate(pool, max_words=max_words, shortlist_bin=sl_bin, want_alignments=True)
    at = 0
    for i, text in enumerate(TEXTS):
        src, tgt, align = blocking[3 * i].split(" ", 2)[2], blocking[3 * i + 1].split(" ", 2)[2], blocking[3 * i + 2]
        n = len(want_segments[i])
        mine, mine_align = outs[at:at + n], st["alignments"][at:at + n]
        at += n
        # 3. target text: every decoded sentence behind the gap in front of its source sentence (Request.cc:147-160)
        fields = src.split()
        assert fields[0] == f"n={n}"
        gaps = [bytes.fromhex(f[1:]).decode() if f != "G-" else "" for f in fields if f.startswith("G")]
        assert len(gaps) == n + 1
        want_text, want_words = "", []
        for s in range(n):
            want_text += gaps[s]
            proto = sp.decode_ids_as_immutable_proto([int(x) for x in mine[s]])
            base = len(want_text.encode())
            cps = [0]
            for ch in proto.text:
                cps.append(cps[-1] + len(ch.encode()))
            want_words.append(",".join(f"{base + cps[p.begin]}:{base + cps[p.end]}" for p in proto.pieces))
            want_text += proto.text
        want_text += gaps[n]
        thex, tann = tgt.split(" ", 1)
        assert (bytes.fromhex(thex).decode() if thex != "-" else "") == want_text
        assert [f[1:] for f in tann.split() if f.startswith("W")] == want_words
        # 4. alignments: the C ABI's rows, sentence by sentence
        got = _floats(align)
        assert len(got) == n
        for s in range(n):
            assert np.array_equal(got[s], mine_align[s].reshape(-1))
    # 5. Async answers like Blocking; a cached answer is the same answer
    for tag in ("async", "cached"):
        mine = [l.split(" ", 1)[1] for l in lines if l.startswith(tag + " ")]
        assert mine == [l.split(" ", 1)[1] for l in blocking[:3]]
    # 6. pivot through the same model: the pivot text is the first pass's target; source side kept; alignments are
    # distributions over the source tokens again
    pivot = [l for l in lines if l.startswith("pivot ")]
    assert len(pivot) == 3 * len(TEXTS)
    for i in range(len(TEXTS)):
        assert pivot[3 * i].split(" ", 2)[2] == blocking[3 * i].split(" ", 2)[2]
        rows = _floats(pivot[3 * i + 2])
        n_src = [len(s) for s in want_segments[i]]
        assert len(rows) == len(n_src)
        for s, row in enumerate(rows):
            m = row.reshape(-1, n_src[s])
            assert np.allclose(m.sum(axis=1), 1.0, atol=1e-3)
    apivot = [l.split(" ", 1)[1] for l in lines if l.startswith("apivot ")]
    assert len(apivot) == 3 and apivot[0] == pivot[0].split(" ", 1)[1]
    if not use_shortlist:  # (with a shortlist the candidate set is the union over the BATCH, Model.cc:116-120: one text alone
        assert apivot == [l.split(" ", 1)[1] for l in pivot[:3]]  # and three texts pooled may legitimately decode differently)
    # 7. a burst of 24 single-line requests through Async: pooled into fewer service calls; without a shortlist every
    # answer equals the same text served alone (batch composition does not matter then)
    burst = dict(kv.split("=") for kv in next(l for l in lines if l.startswith("burst ")).split()[1:])
    assert int(burst["requests"]) == 24 and int(burst["calls"]) < 24
    if not use_shortlist:
        assert int(burst["equal"]) == 24
    assert lines[-1] == "html refused"
