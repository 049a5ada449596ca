"""Synthetic assets for tests and bench: a seeded random-init model in the
marian binary v1 format, a binary lexical shortlist, and token batches.

Formats follow the reference loader (slimt/Io.cc:114-273, slimt/Io.hh:19-29;
shortlist: slimt/Shortlist.hh:78-85, slimt/Shortlist.cc:41-98) and are
documented in SURVEY.md appendices B and C.  There is no network for real
checkpoints, so these stand in for browsermt's `model.intgemm.alphas.bin` and
`lex.s2t.bin`.  Nothing here is on the product path.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Dict, List, Tuple

import numpy as np

TYPE_F32 = 0x0404
TYPE_I8 = 0x0101
TYPE_IG8 = 0x4101
SHORTLIST_MAGIC = 0xF11A48D5013417F5
EOS_ID = 0
PAD_ID = 0


@dataclasses.dataclass
class ModelDims:
    emb: int = 256
    ffn: int = 1536
    heads: int = 8
    vocab: int = 32000
    enc_layers: int = 6
    dec_layers: int = 2


TINY = ModelDims()
BASE = ModelDims(emb=512, ffn=2048)


def _quantize_weight(w: np.ndarray) -> Tuple[np.ndarray, np.float32]:
    """f32 [N,K] -> int8 [N,K] + b_quant = 127/max|w| (never -128)."""
    bq = np.float32(127.0) / np.float32(np.abs(w).max())
    q = np.clip(np.rint(w.astype(np.float32) * bq), -127, 127).astype(np.int8)
    return q, np.float32(bq)


def make_params(dims: ModelDims = TINY, seed: int = 1234, attn_sharpness: float = 1.5,
                eos_bias: float = 2.0, dec_mix: float = 1.0, rnn_scale: float = 3.0, dec_ffn_scale: float = 3.0) -> Dict[str, Tuple[int, Tuple[int, ...], bytes]]:
    """Returns name -> (type, shape, blob).  ig8 blobs hold B^T ([out][in]
    int8) followed by the f32 multiplier, as the reference expects."""
    rng = np.random.RandomState(seed)
    E, F, V = dims.emb, dims.ffn, dims.vocab
    items: Dict[str, Tuple[int, Tuple[int, ...], bytes]] = {}

    def f32(name, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        items[name] = (TYPE_F32, tuple(arr.shape), arr.tobytes())

    def ig8(name, fan_in, fan_out, scale=1.0):
        w = rng.standard_normal((fan_out, fan_in)).astype(np.float32) * np.float32(scale / np.sqrt(fan_in))
        q, bq = _quantize_weight(w)
        items[name] = (TYPE_IG8, (fan_in, fan_out), q.tobytes() + struct.pack("<f", bq))

    def quant(name, alpha):
        f32(name, np.array([127.0 / alpha], dtype=np.float32))

    def ln(prefix):
        f32(prefix + "_ln_scale", 1.0 + 0.1 * rng.standard_normal((1, E)))
        f32(prefix + "_ln_bias", 0.1 * rng.standard_normal((1, E)))

    def attention(prefix, kind):
        for s in "qkvo":
            scale = attn_sharpness if s in "qk" else (dec_mix if kind == "context" else 1.0)
            ig8(f"{prefix}_{kind}_W{s}", E, E, scale)
            f32(f"{prefix}_{kind}_b{s}", 0.05 * rng.standard_normal((1, E)))
            quant(f"{prefix}_{kind}_W{s}_QuantMultA", 4.0 if s == "o" else 6.0)
        ln(f"{prefix}_{kind}_Wo")

    def ffn(prefix, w2_scale=1.0):
        ig8(f"{prefix}_ffn_W1", E, F)
        f32(f"{prefix}_ffn_b1", 0.05 * rng.standard_normal((1, F)))
        quant(f"{prefix}_ffn_W1_QuantMultA", 6.0)
        ig8(f"{prefix}_ffn_W2", F, E, w2_scale)
        f32(f"{prefix}_ffn_b2", 0.05 * rng.standard_normal((1, E)))
        quant(f"{prefix}_ffn_W2_QuantMultA", 6.0)
        ln(f"{prefix}_ffn_ffn")

    # Embedding [V,E]: rows ~ N(0, 1/sqrt(E)) so that emb*sqrt(E) ~ N(0,1).
    wemb = rng.standard_normal((V, E)).astype(np.float32) / np.float32(np.sqrt(E))
    q, bq = _quantize_weight(wemb)
    items["Wemb"] = (TYPE_IG8, (V, E), q.tobytes() + struct.pack("<f", bq))
    quant("none_QuantMultA", 6.0)
    out_b = 0.1 * rng.standard_normal((1, V))
    out_b[0, EOS_ID] = eos_bias
    f32("decoder_ff_logit_out_b", out_b)

    for i in range(1, dims.enc_layers + 1):
        p = f"encoder_l{i}"
        attention(p, "self")
        ffn(p)
    for j in range(1, dims.dec_layers + 1):
        p = f"decoder_l{j}"
        ig8(f"{p}_rnn_W", E, E, rnn_scale)
        quant(f"{p}_rnn_W_QuantMultA", 6.0)
        ig8(f"{p}_rnn_Wf", E, E)
        f32(f"{p}_rnn_bf", 0.05 * rng.standard_normal((1, E)))
        quant(f"{p}_rnn_Wf_QuantMultA", 6.0)
        ln(f"{p}_rnn_ffn")
        attention(p, "context")
        ffn(p, dec_ffn_scale)
    return items


def write_model(path: str, items: Dict[str, Tuple[int, Tuple[int, ...], bytes]]) -> None:
    """Marian binary v1: u64 version, u64 n, n headers, names, shapes, pad, blobs."""
    names = list(items.keys())
    headers, name_blob, shape_blob, data_blob = b"", b"", b"", b""
    for name in names:
        typ, shape, blob = items[name]
        blob = blob + b"\0" * ((-len(blob)) % 256)
        nm = name.encode() + b"\0"
        headers += struct.pack("<QQQQ", len(nm), typ, len(shape), len(blob))
        name_blob += nm
        shape_blob += struct.pack(f"<{len(shape)}i", *shape)
        data_blob += blob
    head = struct.pack("<QQ", 1, len(names)) + headers + name_blob + shape_blob
    pad = (-(len(head) + 8)) % 256
    with open(path, "wb") as f:
        f.write(head + struct.pack("<Q", pad) + b"\0" * pad + data_blob)


def read_model(path: str) -> Dict[str, Tuple[int, Tuple[int, ...], np.ndarray]]:
    """Parses the same format back (used by the numpy oracle and tests)."""
    raw = open(path, "rb").read()
    version, n = struct.unpack_from("<QQ", raw, 0)
    assert version == 1
    off = 16
    hdrs = [struct.unpack_from("<QQQQ", raw, off + 32 * i) for i in range(n)]
    off += 32 * n
    names = []
    for h in hdrs:
        names.append(raw[off:off + h[0] - 1].decode())
        off += h[0]
    shapes = []
    for h in hdrs:
        shapes.append(struct.unpack_from(f"<{h[2]}i", raw, off))
        off += 4 * h[2]
    (pad,) = struct.unpack_from("<Q", raw, off)
    off += 8 + pad
    out = {}
    for name, h, shape in zip(names, hdrs, shapes):
        out[name] = (h[1], tuple(shape), np.frombuffer(raw, dtype=np.uint8, count=h[3], offset=off))
        off += h[3]
    return out


def make_shortlist(vocab: int = 32000, frequent: int = 100, best: int = 100, seed: int = 7,
                   spread: int = 2000) -> Tuple[int, np.ndarray, np.ndarray]:
    """Per-source-word candidate lists: `best` seeded targets each.  Real
    lexical shortlists are topical (candidates of nearby words overlap), so
    targets are drawn from a window of `spread` ids around the source id."""
    rng = np.random.RandomState(seed)
    offsets = np.arange(vocab + 1, dtype=np.uint64) * np.uint64(best)
    centre = np.arange(vocab, dtype=np.int64)[:, None]
    cand = (centre + rng.randint(-spread // 2, spread // 2, size=(vocab, best))) % vocab
    cand = np.sort(cand, axis=1).astype(np.uint32)
    return frequent, offsets, cand.reshape(-1)


def shortlist_checksum(body: bytes) -> int:
    """hash_bytes<uint64_t> (slimt/Utils.hh:46-68) over everything after the header's magic and checksum words:
    boost-style hash_combine with std::hash<uint64_t>, which is the identity in libstdc++."""
    seed, mask = 0, (1 << 64) - 1
    for (w,) in struct.iter_unpack("<Q", body[:len(body) // 8 * 8]):
        seed ^= (w + 0x9e3779b9 + ((seed << 6) & mask) + (seed >> 2)) & mask
    return seed


def write_shortlist(path: str, frequent: int, offsets: np.ndarray, lists: np.ndarray, best: int = 100,
                    checksum: bool = False) -> None:
    """`checksum`: compute the header's checksum (a Python loop over the image: small images only); otherwise it is
    written as 0, which the reference's default check = false never looks at (Shortlist.hh:52)."""
    body = struct.pack("<QQQQ", frequent, best, len(offsets), len(lists)) + offsets.astype("<u8").tobytes() + lists.astype("<u4").tobytes()
    header = struct.pack("<QQ", SHORTLIST_MAGIC, shortlist_checksum(body) if checksum else 0)
    with open(path, "wb") as f:
        f.write(header + body)


def make_sentences(n: int, length, vocab: int = 32000, seed: int = 99) -> List[np.ndarray]:
    """n sentences of ids ~ U{1..vocab-1} ending in EOS (=0).  `length` is an
    int or an inclusive (lo, hi) range for the mixed-length sweep."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        L = length if isinstance(length, int) else int(rng.randint(length[0], length[1] + 1))
        s = rng.randint(1, vocab, size=L).astype(np.uint32)
        s[-1] = EOS_ID
        out.append(s)
    return out


def pack_batch(sentences: List[np.ndarray], limit_factor: float = 1.5,
               shortlist: np.ndarray | None = None) -> bytes:
    """BatchSpec record read by oracle/ref_driver.cc (read_batch)."""
    B = len(sentences)
    T = max(len(s) for s in sentences)
    tok = np.full((B, T), PAD_ID, dtype=np.uint32)
    lens = np.zeros(B, dtype=np.uint32)
    for i, s in enumerate(sentences):
        tok[i, :len(s)] = s
        lens[i] = len(s)
    sl = np.zeros(0, dtype=np.uint32) if shortlist is None else np.asarray(shortlist, dtype=np.uint32)
    return struct.pack("<IIfI", B, T, limit_factor, len(sl)) + lens.tobytes() + tok.tobytes() + sl.tobytes()
