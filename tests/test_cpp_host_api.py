"""The C++ host layer (include/slimt_b200.hh): slimt's Tensor / qmm:: / Input / Model / Blocking / Async surface
over the C ABI.  CPU: the header compiles, links against the library and fails loudly without a GPU.  GPU: its
results equal the ctypes path (which the other GPU suites pin bit-exactly to the oracle)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import sb_testutil as util
from slimt_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_api_test")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "slimt_b200", "csrc"), "host_test"])
    assert os.path.exists(BIN)


def _write_sentences(path, sents):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(sents)))
        for s in sents:
            f.write(struct.pack("<I", len(s)) + np.asarray(s, dtype=np.uint32).tobytes())


def _read_sentences(buf, pos):
    n, = struct.unpack_from("<I", buf, pos)
    pos += 4
    out = []
    for _ in range(n):
        ln, = struct.unpack_from("<I", buf, pos)
        pos += 4
        out.append(np.frombuffer(buf, dtype=np.uint32, count=ln, offset=pos).tolist())
        pos += 4 * ln
    return out, pos


def test_host_layer_builds_and_refuses_to_run_without_gpu(tiny_model, tmp_path):
    _build()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    except ImportError:
        pass
    sents = synth.make_sentences(3, (2, 6), seed=5)
    _write_sentences(tmp_path / "s.u32", sents)
    r = subprocess.run([BIN, tiny_model[0], "-", str(tmp_path / "s.u32"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr


LOGIC = os.path.join(ROOT, "tests", "cpp", "host_logic_test")


def _wrap_ref(words, wrap_length, eos=0):
    """TextProcessor::wrap (slimt/TextProcessor.cc:123-157): segments of wrap_length - 1 words, each with its own EOS."""
    step = wrap_length - 1
    return [words[o:o + step] + [eos] for o in range(0, len(words), step)]


def _cache_key_ref(model_id, words):
    """cache_key (slimt/Request.cc:20-26) over hash_combine (Utils.hh:47-57); std::hash<size_t> is the identity."""
    mask = (1 << 64) - 1
    seed = model_id
    for w in words:
        seed ^= (w + 0x9e3779b9 + ((seed << 6) & mask) + (seed >> 2)) & mask
    return seed


@pytest.mark.parametrize("length,wrap_length", [(12, 6), (5, 6), (6, 6), (1, 2), (300, 128), (255, 128), (10, 1)])
def test_host_wrap_and_cache_key(length, wrap_length):
    """CPU-only pieces of the C++ host layer: wrapping of long sentences on word ids, the cache key, AtomicCache."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "slimt_b200", "csrc"), "host_logic"])
    out = subprocess.run([LOGIC, str(length), str(wrap_length)], capture_output=True, text=True, check=True).stdout.split("\n")
    words = list(range(1, length + 1))
    segs = [list(map(int, l.split()[1:])) for l in out if l.startswith("seg ")]
    if wrap_length < 2 or length + 1 <= wrap_length:  # fits (EOS included): passed through untouched
        assert segs == [words + [0]]
    else:
        assert segs == _wrap_ref(words, wrap_length)
        assert all(len(s) <= wrap_length for s in segs)
    assert int(next(l for l in out if l.startswith("segments ")).split()[1]) == len(segs)
    assert int(next(l for l in out if l.startswith("key ")).split()[1]) == _cache_key_ref(3, words + [0])
    assert "cache ok" in out and "key0 0" in out and "make_cache 1 1" in out


@pytest.mark.gpu
def test_host_layer_matches_ctypes_path(gpu_ctx, tiny_model, shortlist_assets, tmp_path):
    from oracle import slimt_oracle as so
    from slimt_b200 import capi
    _build()
    path, items = tiny_model
    sl_path, (fr, offs, lists) = shortlist_assets
    sents = synth.make_sentences(20, (2, 12), seed=91)
    _write_sentences(tmp_path / "s.u32", sents)
    subprocess.run([BIN, path, sl_path, str(tmp_path / "s.u32"), str(tmp_path / "o.bin")], check=True)
    buf = open(tmp_path / "o.bin", "rb").read()

    # 1. qmm::affine on the same deterministic operands
    # same float32 arithmetic as tests/cpp/host_api_test.cc (0.01f * float(i % 97 - 48), 0.1f * float(i))
    x = (np.float32(0.01) * ((np.arange(4 * 64) % 97) - 48).astype(np.float32)).reshape(4, 64)
    W = (((np.arange(64 * 16) * 37) % 255) - 127).astype(np.int8).reshape(16, 64)
    b = np.float32(0.1) * np.arange(16).astype(np.float32)
    y_ref = so.affine(x, W, b, float(np.float32(127.0) / np.float32(0.5)), float(np.float32(127.0) / np.float32(2.0)))
    y_ref = y_ref[0] if isinstance(y_ref, tuple) else y_ref
    y = np.frombuffer(buf, dtype=np.float32, count=64).reshape(4, 16)
    assert np.array_equal(y, y_ref)
    pos = 256

    # 2. Model::forward == oracle on the same single batch (shortlist = union over the batch)
    fwd, pos = _read_sentences(buf, pos)
    tokens, lengths = util.pad_batch(sents)
    sl = so.shortlist_generate(np.concatenate(sents), fr, offs, lists, 32000)
    ref = so.Oracle(items).forward(tokens, lengths, shortlist=sl, keep=True)
    assert fwd == ref["sentences"]
    n_align, = struct.unpack_from("<I", buf, pos)
    pos += 4
    assert n_align == len(ref["sentences"][0])
    for s in range(n_align):
        row = np.frombuffer(buf, dtype=np.float32, count=int(lengths[0]), offset=pos)
        pos += 4 * int(lengths[0])
        assert np.array_equal(row, ref["attn"][s][0, 0, 0, :lengths[0]])

    # 3. Blocking::translate == the ctypes translate with the same max_words (tokens and Response.alignments)
    m = capi.Model(gpu_ctx, open(path, "rb").read())
    sl_bin = open(sl_path, "rb").read()
    outs, st = m.translate(sents, max_words=96, shortlist_bin=sl_bin, want_alignments=True)
    blocking, pos = _read_sentences(buf, pos)
    assert blocking == [o.tolist() for o in outs]
    for i, s in enumerate(sents):
        rows, = struct.unpack_from("<I", buf, pos)
        pos += 4
        assert rows == len(outs[i])
        a = np.frombuffer(buf, dtype=np.float32, count=rows * len(s), offset=pos).reshape(rows, len(s))
        pos += 4 * rows * len(s)
        assert np.array_equal(a, st["alignments"][i])
    # 4. Async over two replicas on one device, twice; Async::pivot == translate(translate(.))
    a1, pos = _read_sentences(buf, pos)
    a2, pos = _read_sentences(buf, pos)
    a3, pos = _read_sentences(buf, pos)
    assert a1 == blocking
    one, _ = m.translate(sents[:1], max_words=96, shortlist_bin=sl_bin)
    assert a2 == [one[0].tolist()]
    second, _ = m.translate(outs, max_words=96, shortlist_bin=sl_bin)
    assert a3 == [o.tolist() for o in second]
    m.close()


def test_host_shortlist_generator_and_batch_plan(tmp_path):
    """slimt::ShortlistGenerator / Shortlist (Shortlist.hh:14-76) and plan_batches (Batcher.cc:95-120) of the C++ layer
    against the oracle's restatements; CPU only."""
    from oracle import slimt_oracle as so
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "slimt_b200", "csrc"), "host_logic"])
    fr, offs, lists = synth.make_shortlist(vocab=200, frequent=10, best=3, seed=5)
    sl_path = str(tmp_path / "lex.bin")
    synth.write_shortlist(sl_path, fr, offs, lists, best=3, checksum=True)  # check = true wants a real checksum
    out = subprocess.run([LOGIC, "9", "128", sl_path, "200"], capture_output=True, text=True, check=True).stdout.split("\n")
    got = list(map(int, next(l for l in out if l.startswith("shortlist")).split()[1:]))
    want = so.shortlist_generate(np.array(list(range(1, 10)) + [0], dtype=np.uint32), fr, offs, lists, 200)
    assert got == list(map(int, want)) and len(got) % 8 == 0
    maps = next(l for l in out if l.startswith("maps")).split()
    assert int(maps[1]) == len(got) // 2 and int(maps[2]) == -1 and int(maps[3]) == got[len(got) // 2]
    # Batcher::generate: ascending length, (n + 1) * longest <= max_words
    lengths = [(i * 7) % 11 + 1 for i in range(10)]
    plan = [(int(l.split(":")[0].split()[1]), list(map(int, l.split(":")[1].split()))) for l in out if l.startswith("batch")]
    seen = []
    for width, ids in plan:
        assert width == max(lengths[i] for i in ids) and len(ids) * width <= 24
        seen += ids
    assert sorted(seen) == list(range(10)) and [lengths[i] for i in seen] == sorted(lengths)
    assert "truncated refused" in out
