"""Shared helpers for the parity tests."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "slimt_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# (M, K, N) of the reference's own int8 tests: gemmology test_multiply.cpp:414-425, intgemm add127_test.cc:283-423
REFERENCE_GEMM_SHAPES = [(8, 256, 256), (8, 2048, 256), (320, 256, 256), (472, 256, 256), (248, 256, 256), (200, 256, 256)]
# hot-path shapes at tiny11 / base
HOTPATH_GEMM_SHAPES = [(96, 256, 1536), (96, 1536, 256), (40, 512, 2048), (40, 2048, 512), (24, 256, 32000)]


def have_ref():
    return os.path.exists(REF_BIN)


def pad_batch(sentences):
    B = len(sentences)
    T = max(len(s) for s in sentences)
    tokens = np.zeros((B, T), dtype=np.uint32)
    lengths = np.zeros(B, dtype=np.uint32)
    for i, s in enumerate(sentences):
        tokens[i, :len(s)] = s
        lengths[i] = len(s)
    return tokens, lengths


def make_qmm_case(seed, M, K, N, scale=1.5, alpha=5.0):
    rng = np.random.RandomState(seed)
    x = (rng.standard_normal((M, K)) * scale).astype(np.float32)
    w = rng.standard_normal((N, K)).astype(np.float32) / np.float32(np.sqrt(K))
    bq = np.float32(127.0) / np.float32(np.abs(w).max())
    Bt = np.clip(np.rint(w * bq), -127, 127).astype(np.int8)
    bias = rng.standard_normal(N).astype(np.float32)
    return x, Bt, bias, float(np.float32(127.0 / alpha)), float(bq)


def ref_qmm(x, Bt, bias, aq, bq, indices=None, env=None, tmpdir="/tmp"):
    """qmm::affine / affine_with_select through the unmodified reference (oracle/_ref)."""
    M, K = x.shape
    N = Bt.shape[0]
    idx = np.zeros(0, np.uint32) if indices is None else np.asarray(indices, dtype=np.uint32)
    cin, cout = os.path.join(tmpdir, f"qmm_{os.getpid()}.in"), os.path.join(tmpdir, f"qmm_{os.getpid()}.out")
    with open(cin, "wb") as f:
        f.write(struct.pack("<IIIIff", M, K, N, len(idx), aq, bq) + x.astype(np.float32).tobytes() + Bt.tobytes()
                + np.asarray(bias, dtype=np.float32).tobytes() + idx.tobytes())
    e = dict(os.environ)
    e.update(env or {})
    subprocess.run([REF_BIN, "qmm", "--case", cin, "--out", cout], check=True, env=e)
    raw = open(cout, "rb").read()
    os.unlink(cin), os.unlink(cout)
    nout = len(idx) if len(idx) else N
    qa_u8 = np.frombuffer(raw[:M * K], dtype=np.uint8).reshape(M, K)
    y = np.frombuffer(raw[M * K:], dtype=np.float32).reshape(M, nout)
    return y, qa_u8


def ref_op(op, dims, arrays, tmpdir="/tmp"):
    d = list(dims) + [0] * (8 - len(dims))
    cin, cout = os.path.join(tmpdir, f"op_{os.getpid()}.in"), os.path.join(tmpdir, f"op_{os.getpid()}.out")
    with open(cin, "wb") as f:
        f.write(struct.pack("<8I", *d) + b"".join(np.ascontiguousarray(a, dtype=np.float32).tobytes() for a in arrays))
    subprocess.run([REF_BIN, "ops", "--op", op, "--in", cin, "--out", cout], check=True)
    out = np.fromfile(cout, dtype=np.float32)
    os.unlink(cin), os.unlink(cout)
    return out


def ref_forward(model_path, sentences, limit=1.5, shortlist=None, forced=None, dump=False, tmpdir="/tmp", env=None):
    """Model::forward through the unmodified reference. Returns dict like the port oracle's."""
    import shutil
    import tempfile
    from slimt_b200 import synth
    work = tempfile.mkdtemp(prefix="ref_fwd_", dir=tmpdir)
    try:
        open(os.path.join(work, "batch.bin"), "wb").write(synth.pack_batch(sentences, limit, shortlist))
        cmd = [REF_BIN, "forward", "--model", model_path, "--batch", os.path.join(work, "batch.bin"), "--out", work]
        if dump:
            cmd.append("--dump")
        if forced is not None:
            np.asarray(forced, dtype=np.uint32).tofile(os.path.join(work, "forced.u32"))
            cmd += ["--force", os.path.join(work, "forced.u32")]
        e = dict(os.environ)
        e.update(env or {})
        subprocess.run(cmd, check=True, capture_output=True, env=e)
        B = len(sentences)
        T = max(len(s) for s in sentences)
        steps = np.fromfile(os.path.join(work, "step_tokens.u32"), dtype=np.uint32).reshape(-1, B)
        flat = np.fromfile(os.path.join(work, "sentences.u32"), dtype=np.uint32)
        sents, p = [], 1
        for _ in range(int(flat[0])):
            n = int(flat[p])
            sents.append(flat[p + 1:p + 1 + n].tolist())
            p += 1 + n
        out = {"step_tokens": steps, "sentences": sents}
        if dump:
            E = np.fromfile(os.path.join(work, "embed.f32"), dtype=np.float32).size // (B * T)
            out["embed"] = np.fromfile(os.path.join(work, "embed.f32"), dtype=np.float32).reshape(B, T, E)
            out["enc"] = [np.fromfile(os.path.join(work, f"enc_l{i}.f32"), dtype=np.float32).reshape(B, T, E) for i in range(1, 7)]
            out["encoder_out"] = out["enc"][-1]
            out["logits"] = [np.fromfile(os.path.join(work, f"logits_{s}.f32"), dtype=np.float32).reshape(B, -1) for s in range(len(steps))]
            out["attn"] = [np.fromfile(os.path.join(work, f"attn_{s}.f32"), dtype=np.float32).reshape(B, 8, 1, T) for s in range(len(steps))]
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def batcher_generate_py(lengths, max_words):
    """Batcher::generate (slimt/Batcher.cc:95-120) restated: ascending length buckets, greedy fill."""
    buckets = {}
    for i, n in enumerate(lengths):
        buckets.setdefault(int(n), []).append(i)
    heads = {k: 0 for k in buckets}
    out = []
    while True:
        batch, width, full = [], 0, False
        for length in sorted(buckets):
            while heads[length] < len(buckets[length]):
                if (len(batch) + 1) * length <= max_words or not batch:
                    batch.append(buckets[length][heads[length]])
                    heads[length] += 1
                    width = max(width, length)
                else:
                    full = True
                    break
            if full:
                break
        if not batch:
            return out
        out.append((batch, width))
