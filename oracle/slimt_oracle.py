"""oracle/slimt_oracle.py -- TEST INFRASTRUCTURE ("port" oracle), not product code.

CPU restatement of slimt's hot path (Model::forward, reference
slimt/Model.cc:187-204 and everything beneath it) on top of the plain-C
kernels in oracle/slimt_oracle.c.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it.

Parity status: PINNED against the unmodified reference compiled in place
(oracle/_ref/slimt_ref, see oracle/Makefile) by tests/test_oracle_vs_ref.py and
against the committed reference-generated vectors under tests/golden/.

The integer GEMM uses float64 BLAS: every partial sum is an integer below
2**53, so the product is exact -- identical to intgemm's AVX512-VNNI path and
to tcgen05 s8 x s8 -> s32.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i8 = ctypes.POINTER(ctypes.c_int8)
c_i32 = ctypes.POINTER(ctypes.c_int32)
c_u32 = ctypes.POINTER(ctypes.c_uint32)
sz = ctypes.c_size_t


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
        _LIB = ctypes.CDLL(path)
        _LIB.so_saturation_count.restype = ctypes.c_uint64
    return _LIB


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------- primitives
def quantize(x: np.ndarray, aq: float) -> np.ndarray:
    """qa = clamp(rne(x*aq), -127, 127) (reference u8 operand = qa + 127)."""
    x = f32c(x)
    q = np.empty(x.shape, dtype=np.int8)
    lib().so_quantize(_p(x, c_f), _p(q, c_i8), ctypes.c_float(aq), sz(x.size))
    return q


def gemm_shifted(qa: np.ndarray, Bt: np.ndarray, exact: bool = True) -> np.ndarray:
    """acc[r,n] = sum_k (qa[r,k]+127) * Bt[n,k]  (int32)."""
    M, K = qa.shape
    N = Bt.shape[0]
    if exact:
        acc = (qa.astype(np.float64) + 127.0) @ Bt.astype(np.float64).T
        return np.rint(acc).astype(np.int64).astype(np.int32)
    out = np.empty((M, N), dtype=np.int32)
    qa = np.ascontiguousarray(qa)
    Bt = np.ascontiguousarray(Bt)
    lib().so_gemm_shifted(_p(qa, c_i8), _p(Bt, c_i8), _p(out, c_i32), sz(M), sz(K), sz(N), 0)
    return out


def gemm_shifted_c(qa: np.ndarray, Bt: np.ndarray, exact: bool = True) -> np.ndarray:
    """Same, through the scalar C loop (slow; pins the float64-BLAS shortcut)."""
    M, K = qa.shape
    N = Bt.shape[0]
    out = np.empty((M, N), dtype=np.int32)
    qa = np.ascontiguousarray(qa)
    Bt = np.ascontiguousarray(Bt)
    lib().so_gemm_shifted(_p(qa, c_i8), _p(Bt, c_i8), _p(out, c_i32), sz(M), sz(K), sz(N), int(exact))
    return out


def saturation_count(qa: np.ndarray, Bt: np.ndarray) -> int:
    M, K = qa.shape
    qa = np.ascontiguousarray(qa)
    Bt = np.ascontiguousarray(Bt)
    return int(lib().so_saturation_count(_p(qa, c_i8), _p(Bt, c_i8), sz(M), sz(K), sz(Bt.shape[0])))


def prepare_bias(Bt: np.ndarray, bias: Optional[np.ndarray], aq: float, bq: float) -> Tuple[np.ndarray, np.ndarray]:
    N, K = Bt.shape
    pb = np.empty(N, dtype=np.float32)
    cs = np.empty(N, dtype=np.int32)
    Bt = np.ascontiguousarray(Bt)
    b = None if bias is None else f32c(bias).reshape(-1)
    lib().so_prepare_bias(_p(Bt, c_i8), _p(b, c_f) if b is not None else None, ctypes.c_float(aq),
                          ctypes.c_float(bq), sz(K), sz(N), _p(pb, c_f), _p(cs, c_i32))
    return pb, cs


def unquantize(acc: np.ndarray, pb: np.ndarray, aq: float, bq: float, fma: bool = False) -> np.ndarray:
    M, N = acc.shape
    y = np.empty((M, N), dtype=np.float32)
    acc = np.ascontiguousarray(acc, dtype=np.int32)
    pb = f32c(pb)
    lib().so_unquantize(_p(acc, c_i32), _p(pb, c_f), ctypes.c_float(aq), ctypes.c_float(bq), sz(M), sz(N),
                        _p(y, c_f), int(fma))
    return y


def affine(x: np.ndarray, Bt: np.ndarray, bias: Optional[np.ndarray], aq: float, bq: float,
           indices: Optional[np.ndarray] = None, exact: bool = True, fma: bool = False,
           want: bool = False):
    """qmm::affine / dot (bias=None) / affine_with_select (indices given).
    Reference: slimt/qmm/Intgemm.inl.cc:8-243 == slimt/qmm/Gemmology.inl.cc:33-281."""
    shp = x.shape
    x2 = f32c(x).reshape(-1, shp[-1])
    qa = quantize(x2, aq)
    pb, _ = prepare_bias(Bt, bias, aq, bq)  # on the FULL B, then gathered (Intgemm.inl.cc:33-66)
    if indices is not None:
        Bt = Bt[np.asarray(indices, dtype=np.int64)]
        pb = pb[np.asarray(indices, dtype=np.int64)]
    acc = gemm_shifted(qa, Bt, exact=exact)
    y = unquantize(acc, pb, aq, bq, fma=fma).reshape(shp[:-1] + (Bt.shape[0],))
    if want:
        return y, qa, acc
    return y


def layer_norm(x: np.ndarray, scale: np.ndarray, bias: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    x = f32c(x)
    cols = x.shape[-1]
    out = np.empty_like(x)
    lib().so_layer_norm(_p(x, c_f), _p(f32c(scale), c_f), _p(f32c(bias), c_f), ctypes.c_float(eps),
                        sz(x.size // cols), sz(cols), _p(out, c_f))
    return out


def softmax(x: np.ndarray) -> np.ndarray:
    x = f32c(x)
    out = np.empty_like(x)
    lib().so_softmax(_p(x, c_f), sz(x.size // x.shape[-1]), sz(x.shape[-1]), _p(out, c_f))
    return out


def sdpa(q: np.ndarray, k: np.ndarray, v: np.ndarray, mask: np.ndarray, heads: int):
    """q [B,Tq,E], k/v [B,Tk,E], mask [B,Tk] additive -> out [B,Tq,E], attn [B,H,Tq,Tk]."""
    q, k, v, mask = f32c(q), f32c(k), f32c(v), f32c(mask)
    B, Tq, E = q.shape
    Tk = k.shape[1]
    out = np.empty_like(q)
    attn = np.empty((B, heads, Tq, Tk), dtype=np.float32)
    lib().so_sdpa(_p(q, c_f), _p(k, c_f), _p(v, c_f), _p(mask, c_f), sz(B), sz(heads), sz(Tq), sz(Tk),
                  sz(E // heads), _p(out, c_f), _p(attn, c_f))
    return out, attn


def highway(x: np.ndarray, y: np.ndarray, g: np.ndarray) -> np.ndarray:
    x, y, g = f32c(x), f32c(y), f32c(g)
    out = np.empty_like(x)
    lib().so_highway(_p(x, c_f), _p(y, c_f), _p(g, c_f), sz(x.size), _p(out, c_f))
    return out


def sinusoid(start: int, T: int, E: int) -> np.ndarray:
    out = np.empty((T, E), dtype=np.float32)
    lib().so_sinusoid(int(start), sz(T), sz(E), _p(out, c_f))
    return out


def argmax_first(logits: np.ndarray) -> np.ndarray:
    logits = f32c(logits)
    rows = logits.size // logits.shape[-1]
    out = np.empty(rows, dtype=np.uint32)
    lib().so_argmax(_p(logits, c_f), sz(rows), sz(logits.shape[-1]), _p(out, c_u32))
    return out


# ---------------------------------------------------------------- shortlist
def shortlist_generate(words: np.ndarray, frequent: int, offsets: np.ndarray, lists: np.ndarray,
                       vocab: int) -> np.ndarray:
    """ShortlistGenerator::generate (slimt/Shortlist.cc:115-175), shared_=false."""
    table = np.zeros(vocab, dtype=bool)
    table[:min(frequent, vocab)] = True
    for w in np.unique(np.asarray(words, dtype=np.int64)):
        table[lists[int(offsets[w]):int(offsets[w + 1])]] = True
    ones = int(table.sum())
    i = frequent
    while i < vocab and ones % 8 != 0:
        if not table[i]:
            table[i] = True
            ones += 1
        i += 1
    return np.nonzero(table)[0].astype(np.uint32)


# ---------------------------------------------------------------- model
class Weights:
    """Parameters by marian name (slimt/Modules.cc:336-405, Transformer.cc:104-118)."""

    def __init__(self, items: Dict[str, Tuple[int, Tuple[int, ...], np.ndarray]]):
        self.f: Dict[str, np.ndarray] = {}
        self.w: Dict[str, Tuple[np.ndarray, float]] = {}
        for name, (typ, shape, raw) in items.items():
            n = int(np.prod(shape))
            if typ == 0x4101:  # intgemm8: [in,out] logical; blob = out rows x in, then f32 multiplier
                K, N = shape
                # "Wemb" is [V,E] stored as V rows of E (Io.cc:183-200); others as N rows of K
                Bt = raw[:n].view(np.int8).reshape((K, N) if name == "Wemb" else (N, K))
                bq = float(raw[n:n + 4].view(np.float32)[0])
                self.w[name] = (Bt, bq)
            elif typ == 0x0404:
                self.f[name] = raw[:4 * n].view(np.float32).reshape(shape)
        # Wemb [V,E] stored as V rows of E: dequantised table for lookups (Io.cc:275-283) and,
        # re-quantised, the output layer B^T (Io.cc:207-224; -128 -> -127).
        q, qm = self.w.pop("Wemb")
        self.emb_q, self.emb_qm = q, qm
        self.emb_f32 = q.astype(np.float32) * (np.float32(1) / np.float32(qm))
        out_q = np.empty(q.shape, dtype=np.int8)
        ef = np.ascontiguousarray(self.emb_f32)
        lib().so_quantize_weight(_p(ef, c_f), _p(out_q, c_i8), ctypes.c_float(qm), sz(ef.size))
        self.w["Wemb_intgemm8"] = (out_q, qm)
        self.E = q.shape[1]
        self.V = q.shape[0]

    def aff(self, wname: str, bname: Optional[str]):
        Bt, bq = self.w[wname]
        b = self.f[bname] if bname else None
        aq = float(self.f[wname + "_QuantMultA"].reshape(-1)[0])
        return Bt, b, aq, bq


class Oracle:
    def __init__(self, items, heads: int = 8, enc_layers: int = 6, dec_layers: int = 2, exact: bool = True,
                 fma: bool = False):
        self.W = Weights(items)
        self.H, self.Le, self.Ld = heads, enc_layers, dec_layers
        self.exact, self.fma = exact, fma
        self.trace: Dict[str, np.ndarray] = {}

    def _affine(self, x, wname, bname, indices=None):
        Bt, b, aq, bq = self.W.aff(wname, bname)
        return affine(x, Bt, b, aq, bq, indices=indices, exact=self.exact, fma=self.fma)

    def embed(self, tokens: np.ndarray) -> np.ndarray:
        """index_select + transform_embedding (Model.cc:195-197; Transformer.cc:24-49)."""
        B, T = tokens.shape
        x = self.W.emb_f32[tokens.astype(np.int64)]
        x = x * np.float32(np.sqrt(np.float32(self.W.E)))
        return (x + sinusoid(0, T, self.W.E)[None]).astype(np.float32)

    def attention(self, prefix: str, q_in, kv_in, mask):
        """Attention::forward (Modules.cc:287-319)."""
        yq = self._affine(q_in, prefix + "_Wq", prefix + "_bq")
        yk = self._affine(kv_in, prefix + "_Wk", prefix + "_bk")
        yv = self._affine(kv_in, prefix + "_Wv", prefix + "_bv")
        out, attn = sdpa(yq, yk, yv, mask, self.H)
        yo = self._affine(out, prefix + "_Wo", prefix + "_bo")
        y = layer_norm(q_in + yo, self.W.f[prefix + "_Wo_ln_scale"], self.W.f[prefix + "_Wo_ln_bias"])
        return y, attn

    def ffn_block(self, prefix: str, x):
        """FFN1 -> relu -> FFN2 -> add -> LN (Modules.cc:251-257, 326-331)."""
        h = self._affine(x, prefix + "_ffn_W1", prefix + "_ffn_b1")
        h = np.maximum(h, np.float32(0))
        y = self._affine(h, prefix + "_ffn_W2", prefix + "_ffn_b2")
        return layer_norm(y + x, self.W.f[prefix + "_ffn_ffn_ln_scale"], self.W.f[prefix + "_ffn_ffn_ln_bias"])

    def encode(self, tokens, mask, keep: bool = False):
        x = self.embed(tokens)
        if keep:
            self.trace["embed"] = x
        for i in range(1, self.Le + 1):
            p = f"encoder_l{i}"
            x, _ = self.attention(p + "_self", x, x, mask)
            x = self.ffn_block(p, x)
            if keep:
                self.trace[f"enc_l{i}"] = x
        return x

    def ssru(self, prefix: str, state, x):
        """SSRU::forward (Modules.cc:190-235)."""
        f = self._affine(x, prefix + "_rnn_Wf", prefix + "_rnn_bf")
        wx = self._affine(x, prefix + "_rnn_W", None)
        c = highway(state, wx, f)
        y = np.maximum(c, np.float32(0))
        h = layer_norm(x + y, self.W.f[prefix + "_rnn_ffn_ln_scale"], self.W.f[prefix + "_rnn_ffn_ln_bias"])
        return h, c

    def step(self, enc, mask, states: List[np.ndarray], prev: Optional[np.ndarray], shortlist):
        """Decoder::step (Transformer.cc:120-183). prev=None is step 0 (zero embedding)."""
        B = enc.shape[0]
        E = self.W.E
        if prev is None:
            emb = np.zeros((B, 1, E), dtype=np.float32)
        else:
            emb = self.W.emb_f32[prev.astype(np.int64)][:, None, :]
        x = emb * np.float32(np.sqrt(np.float32(E)))
        x = (x + sinusoid(0, 1, E)[None]).astype(np.float32)  # position 0 at every step (quirk Q1)
        attn = None
        for j in range(1, self.Ld + 1):
            p = f"decoder_l{j}"
            h, states[j - 1] = self.ssru(p, states[j - 1], x)
            y, attn = self.attention(p + "_context", h, enc, mask)
            x = self.ffn_block(p, y)
        # output layer: W = Wemb_intgemm8, a_quant = none_QuantMultA (Transformer.cc:104-112)
        Bt, bq = self.W.w["Wemb_intgemm8"]
        b = self.W.f["decoder_ff_logit_out_b"]
        aq = float(self.W.f["none_QuantMultA"].reshape(-1)[0])
        logits = affine(x, Bt, b, aq, bq, indices=shortlist, exact=self.exact, fma=self.fma)
        return logits, attn, x

    def forward(self, tokens: np.ndarray, lengths: np.ndarray, limit_factor: float = 1.5,
                shortlist: Optional[np.ndarray] = None, forced: Optional[np.ndarray] = None,
                keep: bool = False):
        """Model::forward + Model::decode (Model.cc:111-204).  Returns dict with
        step_tokens [steps,B], sentences (recorded until EOS), and traces."""
        B, T = tokens.shape
        mask01 = (np.arange(T)[None, :] < np.asarray(lengths)[:, None]).astype(np.float32)
        mask = ((np.float32(1) - mask01) * np.float32(-99999999.0)).astype(np.float32)  # Input.cc:49-63
        enc = self.encode(tokens, mask, keep=keep)
        states = [np.zeros((B, 1, self.W.E), dtype=np.float32) for _ in range(self.Ld)]
        complete = np.zeros(B, dtype=bool)
        sentences: List[List[int]] = [[] for _ in range(B)]
        step_tokens, logits_all, attn_all = [], [], []
        prev = None
        max_len = int(np.float32(limit_factor) * np.float32(T))
        remaining = B
        i = 0
        while i < max_len and remaining > 0:
            logits, attn, _ = self.step(enc, mask, states, prev, shortlist)
            idx = argmax_first(logits.reshape(B, -1))
            words = shortlist[idx] if shortlist is not None else idx
            if keep:
                logits_all.append(logits.reshape(B, -1))
                attn_all.append(attn)
            step_tokens.append(words.astype(np.uint32))
            for b in range(B):
                if not complete[b]:
                    complete[b] = words[b] == 0
                    sentences[b].append(int(words[b]))
            remaining = B - int(complete.sum())
            prev = words if forced is None else forced[i]
            i += 1
        return {"step_tokens": np.stack(step_tokens), "sentences": sentences, "encoder_out": enc,
                "logits": logits_all, "attn": attn_all, "mask": mask}
