// extern "C" surface declared in include/slimt_b200.h.
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/slimt_b200.h"
#include "engine.cuh"
#include "service.cuh"

struct slimt_b200_ctx {
  sb::Context c;
};
struct slimt_b200_model {
  sb::Model m;
};

extern "C" {

const char* slimt_b200_last_error(void) { return sb::last_error(); }
const char* slimt_b200_version(void) { return "slimt_b200 0.1 (sm_100a, tcgen05 kind::i8)"; }

int slimt_b200_ctx_create(int device, slimt_b200_ctx** out) {
  auto* ctx = new (std::nothrow) slimt_b200_ctx();
  if (!ctx) return 1;
  if (ctx->c.init(device)) {
    delete ctx;
    return 1;
  }
  *out = ctx;
  return 0;
}

void slimt_b200_ctx_destroy(slimt_b200_ctx* ctx) {
  if (!ctx) return;
  ctx->c.destroy();
  delete ctx;
}

int slimt_b200_ctx_synchronize(slimt_b200_ctx* ctx) {
  SB_CUDA(cudaSetDevice(ctx->c.device));
  SB_CUDA(cudaStreamSynchronize(ctx->c.stream));
  return 0;
}

void* slimt_b200_dev_alloc(slimt_b200_ctx* ctx, size_t bytes) {
  cudaSetDevice(ctx->c.device);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    sb::set_error("cudaMalloc failed for " + std::to_string(bytes) + " bytes");
    return nullptr;
  }
  return p;
}
void slimt_b200_dev_free(slimt_b200_ctx* ctx, void* p) {
  cudaSetDevice(ctx->c.device);
  cudaFree(p);
}
int slimt_b200_memcpy_h2d(slimt_b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
  SB_CUDA(cudaSetDevice(ctx->c.device));
  SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->c.stream));
  SB_CUDA(cudaStreamSynchronize(ctx->c.stream));
  ctx->c.h2d_bytes += bytes;
  return 0;
}
int slimt_b200_memcpy_d2h(slimt_b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
  SB_CUDA(cudaSetDevice(ctx->c.device));
  SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->c.stream));
  SB_CUDA(cudaStreamSynchronize(ctx->c.stream));
  ctx->c.d2h_bytes += bytes;
  return 0;
}
int slimt_b200_timer_start(slimt_b200_ctx* ctx) {
  SB_CUDA(cudaSetDevice(ctx->c.device));
  SB_CUDA(cudaEventRecord(ctx->c.ev0, ctx->c.stream));
  return 0;
}
int slimt_b200_timer_stop(slimt_b200_ctx* ctx, double* ms) {
  SB_CUDA(cudaSetDevice(ctx->c.device));
  SB_CUDA(cudaEventRecord(ctx->c.ev1, ctx->c.stream));
  SB_CUDA(cudaEventSynchronize(ctx->c.ev1));
  float t = 0;
  SB_CUDA(cudaEventElapsedTime(&t, ctx->c.ev0, ctx->c.ev1));
  *ms = t;
  return 0;
}
int slimt_b200_flush_l2(slimt_b200_ctx* ctx, size_t bytes) {
  sb::Context& c = ctx->c;
  SB_CUDA(cudaSetDevice(c.device));
  if (bytes > c.flush_bytes) {
    if (c.flush_buf) SB_CUDA(cudaFree(c.flush_buf));
    c.flush_buf = nullptr;
    SB_CUDA(cudaMalloc(&c.flush_buf, bytes));
    c.flush_bytes = bytes;
  }
  SB_CUDA(cudaMemsetAsync(c.flush_buf, 1, bytes, c.stream));
  return 0;
}

void slimt_b200_qmm_prepare_weight_quantized_transposed(const int8_t* input, int8_t* output, size_t rows, size_t cols) {
  // Already B^T row-major [cols][rows]: tcgen05 consumes it as a K-major operand unchanged.
  memcpy(output, input, rows * cols);
}

void slimt_b200_qmm_prepare_weight_transposed(const float* weights, int8_t* prepared, float quantization_multiplier,
                                              size_t cols, size_t rows) {
  sb::host_quantize(weights, prepared, quantization_multiplier, cols * rows);
}

int slimt_b200_qmm_affine(slimt_b200_ctx* ctx, const float* x, size_t M, size_t K, const int8_t* W, size_t N,
                          const float* bias, float a_quant, float b_quant, const uint32_t* indices, size_t n_indices,
                          float* y) {
  return sb::qmm_affine_host(ctx->c, x, M, K, W, N, bias, a_quant, b_quant, indices, n_indices, y, nullptr, nullptr);
}

int slimt_b200_qmm_affine_debug(slimt_b200_ctx* ctx, const float* x, size_t M, size_t K, const int8_t* W, size_t N,
                                const float* bias, float a_quant, float b_quant, const uint32_t* indices,
                                size_t n_indices, float* y, int8_t* qa_out, int32_t* acc_out) {
  return sb::qmm_affine_host(ctx->c, x, M, K, W, N, bias, a_quant, b_quant, indices, n_indices, y, qa_out, acc_out);
}

int slimt_b200_model_create(slimt_b200_ctx* ctx, const void* model_bin, size_t bytes,
                            const slimt_b200_model_config* config, slimt_b200_model** out) {
  if (config->feed_forward_depth != 2) {
    sb::set_error("feed_forward_depth must be 2");
    return 1;
  }
  auto* m = new (std::nothrow) slimt_b200_model();
  if (!m) return 1;
  if (m->m.load(&ctx->c, model_bin, bytes, config->encoder_layers, config->decoder_layers, config->num_heads)) {
    m->m.destroy();
    delete m;
    return 1;
  }
  *out = m;
  return 0;
}

void slimt_b200_model_destroy(slimt_b200_model* model) {
  if (!model) return;
  model->m.destroy();
  delete model;
}

int slimt_b200_model_dims(const slimt_b200_model* model, int32_t* emb, int32_t* ffn, int32_t* vocab) {
  *emb = model->m.E, *ffn = model->m.F, *vocab = model->m.V;
  return 0;
}

int slimt_b200_model_forward(slimt_b200_model* model, slimt_b200_forward_io* io) {
  sb::ForwardArgs a;
  a.tokens = io->tokens, a.lengths = io->lengths, a.B = io->batch, a.T = io->seq;
  a.limit_factor = io->limit_factor;
  a.shortlist = io->shortlist, a.n_shortlist = io->n_shortlist;
  a.forced = io->forced, a.device_io = io->device_io != 0;
  a.step_tokens = io->step_tokens;
  a.encoder_out = io->encoder_out, a.logits = io->logits, a.alignment = io->alignment;
  int rc = sb::model_forward(model->m, a);
  io->steps = a.steps;
  io->target_tokens = a.target_tokens;
  return rc;
}

namespace {
struct LazyShortlist {
  const sb::ShortlistGenerator* gen;
  const std::vector<uint32_t>* words;
  size_t vocab;
  std::vector<uint32_t> out;
};
int lazy_shortlist_cb(void* user, const uint32_t** words, size_t* n) {
  auto* l = static_cast<LazyShortlist*>(user);
  l->out = l->gen->generate(l->words->data(), l->words->size(), l->vocab);
  *words = l->out.data();
  *n = l->out.size();
  return 0;
}
}  // namespace

int slimt_b200_translate(slimt_b200_model* model, slimt_b200_translate_io* io) {
  sb::Model& m = model->m;
  sb::Context& c = *m.ctx;
  const uint64_t l0 = c.launches, h0 = c.h2d_bytes, d0 = c.d2h_bytes;
  sb::ShortlistGenerator gen;
  const bool use_sl = io->shortlist_bin != nullptr && io->shortlist_bytes > 0;
  if (use_sl && gen.load(io->shortlist_bin, io->shortlist_bytes)) return 1;

  sb::Batcher batcher(io->max_words);
  for (size_t i = 0; i < io->n_sentences; i++) batcher.enqueue(i, io->offsets[i + 1] - io->offsets[i]);

  // record() (Model.cc:127-137) keeps each sentence's tokens up to and including its first EOS.  Each batch's kept
  // tokens are packed into one block per batch (no per-sentence containers); the ragged output is written once every
  // length is known.  The padded batch and the step-token matrix travel through one pinned staging block.
  struct Done {
    std::vector<size_t> ids;
    std::vector<uint32_t> kept;  // the batch's sentences back to back, in batch row order
  };
  std::vector<Done> done;
  std::vector<uint32_t> out_len(io->n_sentences, 0);
  // staging layout: [padded tokens + lengths of the current batch][its step tokens]
  size_t in_bytes = 0, steps_bytes = 0;
  {
    sb::Batcher plan(io->max_words);
    for (size_t i = 0; i < io->n_sentences; i++) plan.enqueue(i, io->offsets[i + 1] - io->offsets[i]);
    for (;;) {
      size_t width = 0;
      std::vector<size_t> b = plan.generate(&width);
      if (b.empty()) break;
      const size_t max_steps = static_cast<size_t>(io->limit_factor * static_cast<float>(width));
      in_bytes = std::max(in_bytes, 4 * (b.size() * width + b.size()));
      steps_bytes = std::max(steps_bytes, 4 * (std::max<size_t>(1, max_steps) + 1) * b.size());
    }
  }
  in_bytes = (in_bytes + 255) & ~size_t(255);
  char* stage = c.staging_reserve(in_bytes + steps_bytes + 256);
  if (!stage) return 1;
  io->target_tokens = 0, io->batches = 0, io->device_ms = 0;
  cudaSetDevice(c.device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0, c.stream);
  for (;;) {
    size_t width = 0;
    std::vector<size_t> batch = batcher.generate(&width);
    if (batch.empty()) break;
    // convert(): Batch -> padded Input (Frontend.cc:30-40; Input.cc:20-47), pad id 0
    const size_t B = batch.size();
    uint32_t* tokens = reinterpret_cast<uint32_t*>(stage);
    uint32_t* lengths = tokens + B * width;
    std::vector<uint32_t> words;
    words.reserve(B * width);
    for (size_t r = 0; r < B; r++) {
      const size_t s = batch[r];
      const size_t len = io->offsets[s + 1] - io->offsets[s];
      memcpy(tokens + r * width, io->tokens + io->offsets[s], 4 * len);
      memset(tokens + r * width + len, 0, 4 * (width - len));
      lengths[r] = static_cast<uint32_t>(len);
      words.insert(words.end(), io->tokens + io->offsets[s], io->tokens + io->offsets[s + 1]);
    }
    // Model::decode builds the candidate set before its first step (Model.cc:116-120); here the host does it
    // while the GPU runs the encoder (the callback fires once the encoder kernels are queued)
    LazyShortlist lazy{&gen, &words, static_cast<size_t>(m.V), {}};
    const size_t max_steps = static_cast<size_t>(io->limit_factor * static_cast<float>(width));
    // one row of up to max_steps tokens per sentence (transposed on the device) followed by the recorded lengths
    const size_t stride = std::max<size_t>(1, max_steps);
    uint32_t* rows = reinterpret_cast<uint32_t*>(stage + in_bytes);
    uint32_t* lens = rows + stride * B;
    sb::ForwardArgs a;
    a.tokens = tokens, a.lengths = lengths, a.B = B, a.T = width;
    a.limit_factor = io->limit_factor;
    if (use_sl) a.shortlist_cb = lazy_shortlist_cb, a.shortlist_user = &lazy;
    a.sentence_tokens = rows, a.row_stride = stride, a.target_lengths = lens;
    if (sb::model_forward(m, a)) {
      cudaEventDestroy(e0), cudaEventDestroy(e1);
      return 1;
    }
    size_t kept_total = 0;
    for (size_t r = 0; r < B; r++) {
      out_len[batch[r]] = lens[r];
      kept_total += lens[r];
    }
    std::vector<uint32_t> kept(kept_total);
    size_t pos = 0;
    for (size_t r = 0; r < B; r++) {
      memcpy(kept.data() + pos, rows + r * stride, 4ul * lens[r]);
      pos += lens[r];
    }
    done.push_back(Done{std::move(batch), std::move(kept)});
    io->target_tokens += a.target_tokens;
    io->batches += 1;
  }
  cudaEventRecord(e1, c.stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  io->device_ms = ms;

  std::vector<uint64_t> out_off(io->n_sentences + 1, 0);
  for (size_t i = 0; i < io->n_sentences; i++) out_off[i + 1] = out_off[i] + out_len[i];
  const uint64_t off = out_off[io->n_sentences];
  if (io->out_tokens && off > io->out_capacity) {
    sb::set_error("out_tokens capacity too small");
    return 1;
  }
  if (io->out_offsets) memcpy(io->out_offsets, out_off.data(), 8 * io->n_sentences);
  if (io->out_tokens) {
    for (const Done& d : done) {
      size_t pos = 0;
      for (size_t s : d.ids) {
        memcpy(io->out_tokens + out_off[s], d.kept.data() + pos, 4ul * out_len[s]);
        pos += out_len[s];
      }
    }
  }
  if (io->out_offsets) io->out_offsets[io->n_sentences] = off;
  io->kernel_launches = c.launches - l0;
  io->h2d_bytes = c.h2d_bytes - h0;
  io->d2h_bytes = c.d2h_bytes - d0;
  return 0;
}

uint64_t slimt_b200_kernel_launches(const slimt_b200_ctx* ctx) { return ctx->c.launches; }

int slimt_b200_shortlist_generate(const void* shortlist_bin, size_t shortlist_bytes, const uint32_t* words,
                                  size_t n_words, size_t vocab, uint32_t* out, size_t out_capacity, size_t* n_out) {
  sb::ShortlistGenerator gen;
  if (gen.load(shortlist_bin, shortlist_bytes)) return 1;
  std::vector<uint32_t> r = gen.generate(words, n_words, vocab);
  *n_out = r.size();
  if (r.size() > out_capacity) {
    sb::set_error("shortlist output capacity too small");
    return 1;
  }
  memcpy(out, r.data(), 4 * r.size());
  return 0;
}

int slimt_b200_batcher_plan(const uint64_t* lengths, size_t n, size_t max_words, uint64_t* batch_ids,
                            uint64_t* batch_offsets, uint64_t* widths, size_t* n_batches) {
  sb::Batcher batcher(max_words);
  for (size_t i = 0; i < n; i++) batcher.enqueue(i, lengths[i]);
  size_t nb = 0, pos = 0;
  batch_offsets[0] = 0;
  for (;;) {
    size_t width = 0;
    std::vector<size_t> b = batcher.generate(&width);
    if (b.empty()) break;
    for (size_t id : b) batch_ids[pos++] = id;
    widths[nb] = width;
    batch_offsets[++nb] = pos;
  }
  *n_batches = nb;
  return 0;
}

int slimt_b200_profile_enable(slimt_b200_ctx* ctx, int on) {
  sb::Context& c = ctx->c;
  SB_CUDA(cudaSetDevice(c.device));
  SB_CUDA(cudaStreamSynchronize(c.stream));
  for (auto& r : c.prof) c.event_pool.push_back(r.e0), c.event_pool.push_back(r.e1);
  c.prof.clear();
  c.profiling = on != 0;
  return 0;
}

int slimt_b200_profile_read(slimt_b200_ctx* ctx, slimt_b200_kernel_stat* out, size_t capacity, size_t* n_out) {
  sb::Context& c = ctx->c;
  SB_CUDA(cudaSetDevice(c.device));
  SB_CUDA(cudaStreamSynchronize(c.stream));
  std::vector<slimt_b200_kernel_stat> agg;
  for (auto& r : c.prof) {
    float ms = 0;
    SB_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    size_t i = 0;
    for (; i < agg.size(); i++)
      if (strcmp(agg[i].name, r.tag) == 0) break;
    if (i == agg.size()) {
      slimt_b200_kernel_stat st;
      memset(&st, 0, sizeof(st));
      strncpy(st.name, r.tag, sizeof(st.name) - 1);
      agg.push_back(st);
    }
    agg[i].launches += 1, agg[i].ms += ms, agg[i].ops += r.ops, agg[i].bytes += r.bytes;
    c.event_pool.push_back(r.e0), c.event_pool.push_back(r.e1);
  }
  c.prof.clear();
  *n_out = agg.size();
  for (size_t i = 0; i < agg.size() && i < capacity; i++) out[i] = agg[i];
  return 0;
}

}  // extern "C"
