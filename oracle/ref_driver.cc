// oracle/ref_driver.cc -- TEST INFRASTRUCTURE, not product code.
//
// A thin driver over the UNMODIFIED reference sources (compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/slimt_ref).  It replays
// slimt::Model::forward / Model::decode (reference slimt/Model.cc:111-204) on
// synthetic token batches through the reference's own Transformer / Encoder /
// Decoder / qmm:: code (intgemm provider, ruy sgemm), bypassing only the text
// front half (Vocabulary/TextProcessor need sentencepiece + PCRE2).
//
// Modes
//   forward  --model M --batch B --out DIR [--dump] [--force F]
//   qmm      --case C --out O
//   bench    --model M --batches BS --workers N [--repeat R] [--tokens-out F]
//            F (first repeat only): u32 n_batches, then per batch u32 B, then per sentence u32 len + u32 words[len]
//   ops      --op layer_norm|softmax|highway|sdpa|sinusoid --in I --out O
//
// All files are raw little-endian; layouts are documented next to each reader.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <optional>
#include <string>
#include <thread>
#include <vector>

#include "intgemm/intgemm.h"
#include "slimt/Input.hh"
#include "slimt/Io.hh"
#include "slimt/Modules.hh"
#include "slimt/QMM.hh"
#include "slimt/Tensor.hh"
#include "slimt/TensorOps.hh"
#include "slimt/Transformer.hh"
#include "slimt/Types.hh"

using namespace slimt;  // NOLINT

namespace slimt {
// Defined (non-static) in reference slimt/Modules.cc:24,88,128 but not declared in Modules.hh.
std::tuple<Tensor, Tensor> scaled_dot_product_attention(const Tensor &q, const Tensor &k, const Tensor &v,
                                                        const Tensor &mask);
Tensor split_heads(const Tensor &x, size_t num_heads);
Tensor join_heads(const Tensor &x);
}  // namespace slimt

namespace {

std::vector<char> slurp(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path.c_str());
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  size_t n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf(n);
  if (fread(buf.data(), 1, n, f) != n) exit(2);
  fclose(f);
  return buf;
}

void spit(const std::string &path, const void *data, size_t bytes) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) {
    fprintf(stderr, "cannot write %s\n", path.c_str());
    exit(2);
  }
  fwrite(data, 1, bytes, f);
  fclose(f);
}

void spit(const std::string &path, const Tensor &t) {
  spit(path, t.data<char>(), t.view().size);
}

// Batch file: u32 B, u32 T, f32 limit_factor, u32 n_shortlist,
//             u32 lengths[B], u32 tokens[B*T] (row-major, padded),
//             u32 shortlist[n_shortlist]
struct BatchSpec {
  uint32_t B = 0, T = 0;
  float limit = 1.5F;
  std::vector<uint32_t> lengths;
  std::vector<uint32_t> tokens;
  std::vector<uint32_t> shortlist;
};

const char *read_batch(const char *p, BatchSpec &b) {
  uint32_t nsl = 0;
  memcpy(&b.B, p, 4), p += 4;
  memcpy(&b.T, p, 4), p += 4;
  memcpy(&b.limit, p, 4), p += 4;
  memcpy(&nsl, p, 4), p += 4;
  b.lengths.resize(b.B);
  memcpy(b.lengths.data(), p, 4UL * b.B), p += 4UL * b.B;
  b.tokens.resize(size_t(b.B) * b.T);
  memcpy(b.tokens.data(), p, 4UL * b.B * b.T), p += 4UL * b.B * b.T;
  b.shortlist.resize(nsl);
  memcpy(b.shortlist.data(), p, 4UL * nsl), p += 4UL * nsl;
  return p;
}

Input make_input(const BatchSpec &b) {
  // Mirrors convert() at reference slimt/Frontend.cc:30-40.
  Input input(b.B, b.T, /*pad_id=*/0, b.limit);
  for (uint32_t i = 0; i < b.B; i++) {
    std::vector<uint32_t> words(b.tokens.begin() + size_t(i) * b.T,
                                b.tokens.begin() + size_t(i) * b.T + b.lengths[i]);
    input.add(words);
  }
  input.finalize();
  return input;
}

// First strict maximum -- same loop as greedy_sample / greedy_sample_from_words
// (reference slimt/Transformer.cc:279-339); restated because the reference
// versions take a sentencepiece-backed Vocabulary only to read its size.
Words argmax_rows(const Tensor &logits, size_t batch, const std::optional<Words> &sl) {
  size_t stride = logits.dim(-1);
  Words out;
  const float *data = logits.data<float>();
  for (size_t i = 0; i < batch; i++) {
    size_t best = 0;
    float bv = data[i * stride];
    for (size_t c = 1; c < stride; c++) {
      float v = data[i * stride + c];
      if (v > bv) bv = v, best = c;
    }
    out.push_back(sl ? (*sl)[best] : Word(best));
  }
  return out;
}

struct ForwardResult {
  size_t steps = 0;
  std::vector<uint32_t> step_tokens;  // [steps][B] raw argmax per step
  std::vector<Words> sentences;       // recorded (stop after EOS)
  size_t target_tokens = 0;
};

// Replays Model::forward + Model::decode (reference slimt/Model.cc:111-204).
ForwardResult run_forward(const Transformer &tf, const BatchSpec &b, const std::string &dump_dir,
                          const std::vector<uint32_t> *forced) {
  Input input = make_input(b);
  const bool dump = !dump_dir.empty();
  Tensor x = index_select(tf.embedding(), input.indices(), "word_embedding");
  transform_embedding(x);
  if (dump) spit(dump_dir + "/embed.f32", x);

  const auto &layers = tf.encoder().encoder();
  for (size_t i = 0; i < layers.size(); i++) {
    auto [y, attn] = layers[i].forward(x, input.mask());
    x = std::move(y);
    if (dump) spit(dump_dir + "/enc_l" + std::to_string(i + 1) + ".f32", x);
  }
  Tensor &encoder_out = x;

  std::optional<Words> indices = std::nullopt;
  if (!b.shortlist.empty()) indices = b.shortlist;

  const uint32_t eos = 0;
  size_t B = b.B;
  std::vector<bool> complete(B, false);
  ForwardResult r;
  r.sentences.resize(B);
  auto record = [&](const Words &step) {
    size_t finished = 0;
    for (size_t i = 0; i < B; i++) {
      if (!complete[i]) {
        complete[i] = (step[i] == eos);
        r.sentences[i].push_back(step[i]);
        r.target_tokens++;
      }
      finished += complete[i] ? 1 : 0;
    }
    return B - finished;
  };

  const Decoder &decoder = tf.decoder();
  Words previous = {};
  std::vector<Tensor> states = decoder.start_states(B);
  size_t max_seq_length = input.limit_factor() * b.T;  // same float->size_t as Model.cc:160
  size_t remaining = B;
  for (size_t i = 0; i < max_seq_length && remaining > 0; i++) {
    auto [logits, attn] = decoder.step(encoder_out, input.mask(), states, previous, indices);
    if (dump) {
      spit(dump_dir + "/logits_" + std::to_string(i) + ".f32", logits);
      spit(dump_dir + "/attn_" + std::to_string(i) + ".f32", attn);
    }
    previous = argmax_rows(logits, B, indices);
    r.step_tokens.insert(r.step_tokens.end(), previous.begin(), previous.end());
    remaining = record(previous);
    r.steps++;
    if (forced) {  // teacher forcing: feed the given token instead of the argmax
      for (size_t j = 0; j < B; j++) previous[j] = (*forced)[i * B + j];
    }
  }
  return r;
}

int mode_forward(int argc, char **argv) {
  std::string model, batch, out, force;
  bool dump = false;
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--model") model = argv[++i];
    else if (a == "--batch") batch = argv[++i];
    else if (a == "--out") out = argv[++i];
    else if (a == "--force") force = argv[++i];
    else if (a == "--dump") dump = true;
  }
  io::MmapFile mm(model);
  Transformer tf(6, 2, 8, 2, View{mm.data(), mm.size()});
  auto raw = slurp(batch);
  BatchSpec b;
  read_batch(raw.data(), b);
  std::vector<uint32_t> forced;
  if (!force.empty()) {
    auto fr = slurp(force);
    forced.resize(fr.size() / 4);
    memcpy(forced.data(), fr.data(), fr.size());
  }
  ForwardResult r = run_forward(tf, b, dump ? out : std::string(), force.empty() ? nullptr : &forced);
  spit(out + "/step_tokens.u32", r.step_tokens.data(), r.step_tokens.size() * 4);
  // sentences: u32 n, then per sentence u32 len + tokens
  std::vector<uint32_t> flat;
  flat.push_back(r.sentences.size());
  for (auto &s : r.sentences) {
    flat.push_back(s.size());
    flat.insert(flat.end(), s.begin(), s.end());
  }
  spit(out + "/sentences.u32", flat.data(), flat.size() * 4);
  printf("{\"steps\": %zu, \"target_tokens\": %zu}\n", r.steps, r.target_tokens);
  return 0;
}

// QMM case file: u32 M, K, N, n_idx; f32 a_quant, b_quant;
//   f32 x[M*K]; i8 Wt[N*K] (stored "transposed": N rows of K); f32 bias[N]; u32 idx[n_idx]
// Output: u8 qa[M*K] (Int8Shift::PrepareA), f32 y[M*Nout]
int mode_qmm(int argc, char **argv) {
  std::string cs, out;
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--case") cs = argv[++i];
    else if (a == "--out") out = argv[++i];
  }
  auto raw = slurp(cs);
  const char *p = raw.data();
  uint32_t M, K, N, nidx;
  float aq, bq;
  memcpy(&M, p, 4), p += 4;
  memcpy(&K, p, 4), p += 4;
  memcpy(&N, p, 4), p += 4;
  memcpy(&nidx, p, 4), p += 4;
  memcpy(&aq, p, 4), p += 4;
  memcpy(&bq, p, 4), p += 4;
  Tensor x(Type::f32, Shape({M, K}), "x");
  memcpy(x.data<char>(), p, 4UL * M * K), p += 4UL * M * K;
  Aligned stored(64, size_t(N) * K);
  memcpy(stored.data(), p, size_t(N) * K), p += size_t(N) * K;
  Tensor bias(Type::f32, Shape({1, N}), "b");
  memcpy(bias.data<char>(), p, 4UL * N), p += 4UL * N;
  std::vector<uint32_t> idx(nidx);
  memcpy(idx.data(), p, 4UL * nidx);

  // Load-time weight prepare exactly as Io.cc:227-242.
  Aligned prepared(64, size_t(N) * K + sizeof(float));
  qmm::prepare_weight_quantized_transposed(reinterpret_cast<int8_t *>(stored.data()),
                                           reinterpret_cast<int8_t *>(prepared.data()), K, N);
  Tensor W;
  W.load(View{prepared.data(), size_t(N) * K}, Type::i8, Shape({K, N}), "W");

  Tensor qa(Type::i8, Shape({M, K}), "qa");
  intgemm::Int8Shift::PrepareA(x.data<float>(), qa.data<int8_t>(), aq, M, K);

  Tensor y = nidx ? qmm::affine_with_select(x, W, bias, aq, bq, idx, "y") : qmm::affine(x, W, bias, aq, bq, "y");
  FILE *f = fopen(out.c_str(), "wb");
  fwrite(qa.data<char>(), 1, size_t(M) * K, f);
  fwrite(y.data<char>(), 1, y.view().size, f);
  fclose(f);
  return 0;
}

// Batches file: u32 n_batches, then n_batches BatchSpec records back to back.
// Workers pull batches from a shared atomic cursor like Async's worker loop
// (reference slimt/Frontend.cc:207-227), all sharing one const Transformer.
int mode_bench(int argc, char **argv) {
  std::string model, batches, tokens_out;
  int workers = 1, repeat = 1;
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--model") model = argv[++i];
    else if (a == "--batches") batches = argv[++i];
    else if (a == "--workers") workers = atoi(argv[++i]);
    else if (a == "--repeat") repeat = atoi(argv[++i]);
    else if (a == "--tokens-out") tokens_out = argv[++i];
  }
  io::MmapFile mm(model);
  Transformer tf(6, 2, 8, 2, View{mm.data(), mm.size()});
  auto raw = slurp(batches);
  const char *p = raw.data();
  uint32_t nb;
  memcpy(&nb, p, 4), p += 4;
  std::vector<BatchSpec> specs(nb);
  for (auto &s : specs) p = read_batch(p, s);
  std::vector<std::vector<Words>> kept(tokens_out.empty() ? 0 : specs.size());

  for (int rep = 0; rep < repeat; rep++) {
    std::atomic<size_t> cursor{0};
    std::atomic<size_t> tokens{0}, src_tokens{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; w++) {
      pool.emplace_back([&]() {
        for (;;) {
          size_t i = cursor.fetch_add(1);
          if (i >= specs.size()) break;
          ForwardResult r = run_forward(tf, specs[i], "", nullptr);
          tokens += r.target_tokens;
          if (rep == 0 && !kept.empty()) kept[i] = std::move(r.sentences);
          for (auto l : specs[i].lengths) src_tokens += l;
        }
      });
    }
    for (auto &t : pool) t.join();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"rep\": %d, \"seconds\": %.6f, \"target_tokens\": %zu, \"source_tokens\": %zu, \"workers\": %d, "
           "\"target_tokens_per_s\": %.3f}\n",
           rep, sec, size_t(tokens), size_t(src_tokens), workers, double(tokens) / sec);
    fflush(stdout);
  }
  if (!tokens_out.empty()) {
    std::vector<uint32_t> flat;
    flat.push_back(uint32_t(kept.size()));
    for (const auto &batch : kept) {
      flat.push_back(uint32_t(batch.size()));
      for (const Words &w : batch) {
        flat.push_back(uint32_t(w.size()));
        flat.insert(flat.end(), w.begin(), w.end());
      }
    }
    spit(tokens_out, flat.data(), 4 * flat.size());
  }
  return 0;
}

// Single f32 ops through the reference's own TensorOps / Modules functions.
// Input file: u32 dims[8] (meaning per op), then f32 payload; output: f32 payload.
int mode_ops(int argc, char **argv) {
  std::string op, in, out;
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--op") op = argv[++i];
    else if (a == "--in") in = argv[++i];
    else if (a == "--out") out = argv[++i];
  }
  auto raw = slurp(in);
  uint32_t d[8];
  memcpy(d, raw.data(), 32);
  const char *p = raw.data() + 32;
  auto take = [&p](Tensor &t) {
    memcpy(t.data<char>(), p, t.view().size);
    p += t.view().size;
  };
  FILE *f = fopen(out.c_str(), "wb");
  auto put = [&f](const Tensor &t) { fwrite(t.data<char>(), 1, t.view().size, f); };
  if (op == "layer_norm") {  // d = rows, cols
    Tensor x(Type::f32, Shape({d[0], d[1]}), "x"), s(Type::f32, Shape({1, d[1]}), "s"), b(Type::f32, Shape({1, d[1]}), "b");
    take(x), take(s), take(b);
    put(layer_norm(x, s, b));
  } else if (op == "softmax") {  // d = rows, cols
    Tensor x(Type::f32, Shape({d[0], d[1]}), "x"), y(Type::f32, Shape({d[0], d[1]}), "y");
    take(x);
    softmax(x.data<float>(), d[0], d[1], y.data<float>());
    put(y);
  } else if (op == "highway") {  // d = n ; highway(x, y, g)
    Tensor x(Type::f32, Shape({d[0]}), "x"), y(Type::f32, Shape({d[0]}), "y"), g(Type::f32, Shape({d[0]}), "g");
    take(x), take(y), take(g);
    put(highway(x, y, g));
  } else if (op == "sdpa") {  // d = B, H, Tq, Tk, dh ; mask is additive [B,Tk]
    uint64_t B = d[0], H = d[1], Tq = d[2], Tk = d[3], E = uint64_t(d[1]) * d[4];
    Tensor q(Type::f32, Shape({B, Tq, E}), "q"), k(Type::f32, Shape({B, Tk, E}), "k"), v(Type::f32, Shape({B, Tk, E}), "v");
    Tensor mask(Type::f32, Shape({B, Tk}), "mask");
    take(q), take(k), take(v), take(mask);
    auto [o, attn] = scaled_dot_product_attention(split_heads(q, H), split_heads(k, H), split_heads(v, H), mask);
    put(join_heads(o));
    put(attn);
  } else if (op == "sinusoid") {  // d = start, T, E
    Tensor y(Type::f32, Shape({d[1], d[2]}), "y");
    sinusoidal_signal(d[0], d[1], d[2], y.data<float>());
    put(y);
  } else {
    fprintf(stderr, "unknown op %s\n", op.c_str());
    return 2;
  }
  fclose(f);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: slimt_ref forward|qmm|bench ...\n");
    return 2;
  }
  std::string mode = argv[1];
  if (mode == "forward") return mode_forward(argc, argv);
  if (mode == "qmm") return mode_qmm(argc, argv);
  if (mode == "bench") return mode_bench(argc, argv);
  if (mode == "ops") return mode_ops(argc, argv);
  fprintf(stderr, "unknown mode %s\n", mode.c_str());
  return 2;
}
