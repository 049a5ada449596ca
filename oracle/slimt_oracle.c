/* oracle/slimt_oracle.c -- TEST INFRASTRUCTURE ("port" oracle), not product code.
 *
 * Plain-C restatement of the arithmetic on slimt's int8 hot path, one function
 * per reference op, each citing the reference file:line it follows.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_ref.py checks every function
 * below against the unmodified reference compiled in place (oracle/_ref), and
 * tests/golden/ holds reference-generated vectors for the GPU box.
 *
 * Compile with -ffp-contract=off and no -march so every f32 op is a single
 * IEEE operation in source order, like the reference's scalar loops built
 * without -mfma.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* ---- quantize A ---------------------------------------------------------
 * intgemm Int8Shift::PrepareA -> QuantizeU (3rd-party/intgemm/intgemm/
 * avx512_gemm.h:256-269; gemmology.h:650-655,823-846): t = x*aq (one f32 mul),
 * cvtps2dq (round-to-nearest-even; NaN or |t| >= 2^31 -> 0x80000000),
 * clamp to [-127,127] (so the overflow sentinel becomes -127), +127 -> u8.
 * We return the signed value qa in [-127,127]; the reference's u8 is qa+127. */
static inline int32_t cvtps2dq(float t) {
  if (!(t == t)) return INT32_MIN;               /* NaN */
  if (t >= 2147483648.0f || t < -2147483648.0f) return INT32_MIN;
  return (int32_t)lrintf(t);                     /* default rounding mode = RNE */
}

void so_quantize(const float *x, int8_t *qa, float aq, size_t n) {
  for (size_t i = 0; i < n; i++) {
    float t = x[i] * aq;
    int32_t v = cvtps2dq(t);
    if (v > 127) v = 127;
    if (v < -127) v = -127;
    qa[i] = (int8_t)v;
  }
}

/* ---- prepare_weight_transposed (Io.cc:207-224 -> intgemm PrepareBTransposed;
 * gemmology.h:1111-1132, 769, 806): quantize f32 -> int8 with -128 banned. */
void so_quantize_weight(const float *w, int8_t *q, float qm, size_t n) {
  so_quantize(w, q, qm, n);
}

/* ---- integer GEMM -------------------------------------------------------
 * Int8Shift::Multiply (intgemm/multiply.h:293-361; avx512vnni_gemm.h:84-122):
 * acc[r][n] = sum_k (qa[r][k]+127) * B[k][n], int32, with B given in the
 * STORED layout Bt[n][k] (file holds B transposed, Appendix B of SURVEY.md).
 * exact=1: VNNI / exact accumulation.  exact=0: maddubs semantics -- adjacent
 * k pairs are summed with int16 saturation first (gemmology.h:576-581). */
void so_gemm_shifted(const int8_t *qa, const int8_t *Bt, int32_t *acc, size_t M, size_t K, size_t N,
                     int exact) {
  for (size_t r = 0; r < M; r++) {
    for (size_t n = 0; n < N; n++) {
      const int8_t *a = qa + r * K;
      const int8_t *b = Bt + n * K;
      int32_t s = 0;
      if (exact) {
        for (size_t k = 0; k < K; k++) s += ((int32_t)a[k] + 127) * (int32_t)b[k];
      } else {
        for (size_t k = 0; k < K; k += 2) {
          int32_t p = ((int32_t)a[k] + 127) * (int32_t)b[k] + ((int32_t)a[k + 1] + 127) * (int32_t)b[k + 1];
          if (p > 32767) p = 32767;
          if (p < -32768) p = -32768;
          s += p;
        }
      }
      acc[r * N + n] = s;
    }
  }
}

/* Count of adjacent-k pairs whose u8*s8 pair sum leaves int16: the
 * "saturation cases" BASELINE.json asks to report separately. */
uint64_t so_saturation_count(const int8_t *qa, const int8_t *Bt, size_t M, size_t K, size_t N) {
  uint64_t c = 0;
  for (size_t r = 0; r < M; r++)
    for (size_t n = 0; n < N; n++)
      for (size_t k = 0; k < K; k += 2) {
        int32_t p = ((int32_t)qa[r * K + k] + 127) * (int32_t)Bt[n * K + k] +
                    ((int32_t)qa[r * K + k + 1] + 127) * (int32_t)Bt[n * K + k + 1];
        c += (p > 32767 || p < -32768);
      }
  return c;
}

/* ---- prepared bias (qmm/Intgemm.inl.cc:112-128; gemmology.h:1274-1318) ---
 * colsum[n] = sum_k B[k][n]; m = (-1*((127/aq)*(127/bq)))/127;
 * pb[n] = float(colsum[n])*m + bias[n]   (bias may be NULL == zeros: dot) */
void so_prepare_bias(const int8_t *Bt, const float *bias, float aq, float bq, size_t K, size_t N,
                     float *pb, int32_t *colsum_out) {
  float a_alpha = 127.0f / aq;
  float b_alpha = 127.0f / bq;
  float m = (-1.0f * (a_alpha * b_alpha)) / 127.0f;
  for (size_t n = 0; n < N; n++) {
    int32_t cs = 0;
    for (size_t k = 0; k < K; k++) cs += Bt[n * K + k];
    if (colsum_out) colsum_out[n] = cs;
    float v = (float)cs * m;
    pb[n] = v + (bias ? bias[n] : 0.0f);
  }
}

/* ---- epilogue UnquantizeAndAddBiasAndWrite (intgemm callbacks; gemmology.h:
 * 969-972,1043-1048): y = float(acc)*um + pb[n], um = 1/(aq*bq).
 * fma=1 evaluates it as one fused multiply-add (what gcc emits inside
 * intgemm's avx512 target functions); fma=0 as mul then add. */
void so_unquantize(const int32_t *acc, const float *pb, float aq, float bq, size_t M, size_t N, float *y,
                   int fma) {
  float um = 1.0f / (aq * bq);
  for (size_t r = 0; r < M; r++)
    for (size_t n = 0; n < N; n++) {
      float a = (float)acc[r * N + n];
      if (fma) {
        y[r * N + n] = fmaf(a, um, pb[n]);
      } else {
        float t = a * um;
        y[r * N + n] = t + pb[n];
      }
    }
}

/* ---- layer_norm (TensorOps.cc:542-580), eps 1e-6 (TensorOps.hh:67-68) ---- */
void so_layer_norm(const float *in, const float *scale, const float *bias, float eps, size_t rows,
                   size_t cols, float *out) {
  for (size_t j = 0; j < rows; j++) {
    const float *x = in + j * cols;
    float *y = out + j * cols;
    float sum = 0.0f;
    for (size_t i = 0; i < cols; i++) sum += x[i];
    float mean = sum / cols;
    float sq = 0.0f;
    for (size_t i = 0; i < cols; i++) {
      float v = x[i] - mean;
      sq += v * v;
    }
    float sigma = sqrtf(sq / cols + eps);
    for (size_t i = 0; i < cols; i++) y[i] = scale[i] * ((x[i] - mean) / sigma) + bias[i];
  }
}

/* ---- softmax (TensorOps.cc:282-315), scalar std::exp ---- */
void so_softmax(const float *logits, size_t rows, size_t cols, float *out) {
  for (size_t i = 0; i < rows; i++) {
    const float *xs = logits + i * cols;
    float mx = -3.402823466e+38f;
    for (size_t j = 0; j < cols; j++) mx = xs[j] > mx ? xs[j] : mx;
    float se = 0.0f;
    for (size_t j = 0; j < cols; j++) se += expf(xs[j] - mx);
    for (size_t j = 0; j < cols; j++) out[i * cols + j] = expf(xs[j] - mx) / se;
  }
}

/* ---- scaled_dot_product_attention (Modules.cc:24-86) on UNSPLIT layouts ----
 * q [B,Tq,H*dh], k,v [B,Tk,H*dh] (split_heads/join_heads, Modules.cc:88-143,
 * are pure permutations, so heads are addressed by stride here);
 * mask [B,Tk] additive (0 / -99999999, Input.cc:49-63).
 * scores = (q.k) * (1/sqrt(dh)) (post-scaled as in TensorOps.cc:436-447),
 * + mask, softmax, out = attn @ v.  attn [B,H,Tq,Tk] is returned when not NULL.
 * Dot products accumulate sequentially in k with separate mul/add; ruy's
 * AVX-512 sgemm uses FMA chains, so agreement with _ref is to ~1e-6, not bits. */
void so_sdpa(const float *q, const float *k, const float *v, const float *mask, size_t B, size_t H,
             size_t Tq, size_t Tk, size_t dh, float *out, float *attn_out) {
  size_t E = H * dh;
  float dk = 1.0f / sqrtf((float)dh);
  float sc[1024];
  float pr[1024];
  for (size_t b = 0; b < B; b++)
    for (size_t h = 0; h < H; h++)
      for (size_t i = 0; i < Tq; i++) {
        const float *qr = q + (b * Tq + i) * E + h * dh;
        for (size_t j = 0; j < Tk; j++) {
          const float *kr = k + (b * Tk + j) * E + h * dh;
          float s = 0.0f;
          for (size_t d = 0; d < dh; d++) s = fmaf(qr[d], kr[d], s);
          s = dk * s;
          sc[j] = s + mask[b * Tk + j];
        }
        so_softmax(sc, 1, Tk, pr);
        if (attn_out) memcpy(attn_out + ((b * H + h) * Tq + i) * Tk, pr, Tk * sizeof(float));
        float *o = out + (b * Tq + i) * E + h * dh;
        for (size_t d = 0; d < dh; d++) {
          float s = 0.0f;
          for (size_t j = 0; j < Tk; j++) s = fmaf(pr[j], v[(b * Tk + j) * E + h * dh + d], s);
          o[d] = s;
        }
      }
}

/* ---- sigmoid / highway (TensorOps.cc:33-36, 662-682) ----
 * out = sg*x + (1-sg)*y with sg = sigmoid(g); SSRU calls highway(c, Wx, f)
 * (Modules.cc:223). */
static inline float so_sigmoid1(float x) {
  return x > 0 ? (1.0f / (1.0f + expf(-x))) : (expf(x) / (1.0f + expf(x)));
}
void so_highway(const float *x, const float *y, const float *g, size_t n, float *out) {
  for (size_t i = 0; i < n; i++) {
    float sg = so_sigmoid1(g[i]);
    float a = sg * x[i];
    float b = (1.0f - sg) * y[i];
    out[i] = a + b;
  }
}

/* ---- sinusoidal_signal (TensorOps.cc:245-265) ---- */
void so_sinusoid(int start, size_t T, size_t E, float *out) {
  float num_timescales = (float)E / 2;
  float inc = logf(10000.0f) / (num_timescales - 1.0f);
  for (size_t p = (size_t)start; p < T + (size_t)start; ++p)
    for (int i = 0; i < num_timescales; ++i) {
      float v = p * expf(i * -inc);
      size_t off = (p - start) * E + i;
      out[off] = sinf(v);
      out[off + (int)num_timescales] = cosf(v);
    }
}

/* ---- greedy argmax, first strict maximum (Transformer.cc:279-339) ---- */
void so_argmax(const float *logits, size_t rows, size_t cols, uint32_t *out) {
  for (size_t i = 0; i < rows; i++) {
    size_t best = 0;
    float bv = logits[i * cols];
    for (size_t c = 1; c < cols; c++) {
      float v = logits[i * cols + c];
      if (v > bv) bv = v, best = c;
    }
    out[i] = (uint32_t)best;
  }
}
