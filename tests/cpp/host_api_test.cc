// Exercises the C++ host layer (include/slimt_b200.hh) the way slimt's own callers use the reference:
// qmm::affine on Tensors, Model::forward on an Input, Blocking::translate and Async::translate on word ids.
// Usage: host_api_test <model.bin> <shortlist.bin|-> <sentences.u32> <out.bin>
//   sentences.u32: u32 n, then per sentence u32 len + words.   out.bin: see write() calls below.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>

#include "slimt_b200.hh"

static std::vector<char> slurp(const char *path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void put(std::ofstream &o, const void *p, size_t n) { o.write(static_cast<const char *>(p), n); }
static void put_sentences(std::ofstream &o, const slimt::Sentences &s) {
  uint32_t n = s.size();
  put(o, &n, 4);
  for (const auto &w : s) {
    uint32_t len = w.size();
    put(o, &len, 4);
    put(o, w.data(), 4 * w.size());
  }
}

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  try {
    std::vector<char> model_bin = slurp(argv[1]);
    std::vector<char> sl_bin = std::string(argv[2]) == "-" ? std::vector<char>() : slurp(argv[2]);
    std::vector<char> raw = slurp(argv[3]);
    if (model_bin.empty() || raw.size() < 4) throw std::runtime_error("cannot read the model or the sentences file");
    const uint32_t *p = reinterpret_cast<const uint32_t *>(raw.data());
    slimt::Sentences sources(*p++);
    size_t longest = 0;
    for (auto &s : sources) {
      uint32_t len = *p++;
      s.assign(p, p + len);
      p += len;
      longest = std::max<size_t>(longest, len);
    }
    std::ofstream out(argv[4], std::ios::binary);

    // 0. host-only pieces: presets (Model.cc:206-245) and the loader's error convention (Io.cc:291-297)
    if (slimt::preset::nano().encoder_layers != 4 || slimt::preset::tiny().encoder_layers != 6 ||
        slimt::preset::base().decoder_layers != 2)
      throw std::runtime_error("preset values");
    bool threw = false;
    try {
      slimt::io::MmapFile missing("/nonexistent/slimt_b200/model.bin");
    } catch (const std::runtime_error &e) {
      threw = std::string(e.what()).find("Failed to open file") == 0;
    }
    if (!threw) throw std::runtime_error("MmapFile did not report a missing file");

    // 1. qmm::affine on Tensors (QMM.hh:48): x [4,64] f32, W {64,16} ig8 as B^T [16][64], bias [1,16]
    slimt::Tensor x(slimt::Type::f32, slimt::Shape({4, 64}), "x");
    slimt::Tensor W(slimt::Type::ig8, slimt::Shape({64, 16}), "W");
    slimt::Tensor b(slimt::Type::f32, slimt::Shape({1, 16}), "b");
    for (size_t i = 0; i < x.size(); i++) x.data<float>()[i] = 0.01f * static_cast<float>(static_cast<int>(i % 97) - 48);
    for (size_t i = 0; i < W.size(); i++) W.data<int8_t>()[i] = static_cast<int8_t>(static_cast<int>((i * 37) % 255) - 127);
    for (size_t i = 0; i < b.size(); i++) b.data<float>()[i] = 0.1f * static_cast<float>(i);
    slimt::Tensor y = slimt::qmm::affine(x, W, b, 127.0f / 0.5f, 127.0f / 2.0f, "y");
    put(out, y.data<float>(), 4 * y.size());

    // 2. Model::forward on one Input holding every sentence (Model.cc:187).  The model comes through the path
    // constructor (Model.hh:53: files mapped with io::MmapFile), a second one through the View constructor.
    slimt::Package<std::string> paths{argv[1], "", std::string(argv[2]) == "-" ? "" : argv[2]};
    auto model = std::make_shared<slimt::Model>(slimt::preset::tiny(), paths);
    {
      slimt::Package<slimt::View> package{{model_bin.data(), model_bin.size()}, {nullptr, 0}, {sl_bin.data(), sl_bin.size()}};
      slimt::Model from_views(slimt::Model::Config{}, package);
      if (from_views.vocabulary_size() != model->vocabulary_size()) throw std::runtime_error("constructors disagree");
    }
    slimt::Input input(sources.size(), longest, 0, 1.5f);
    for (const auto &s : sources) input.add(s);
    input.finalize();
    slimt::Histories histories = model->forward(input);
    slimt::Sentences forward_targets;
    for (const auto &h : histories) forward_targets.push_back(h->target);
    put_sentences(out, forward_targets);
    uint32_t n_align = histories.empty() ? 0 : histories[0]->alignment.size();
    put(out, &n_align, 4);
    for (uint32_t s = 0; s < n_align; s++) put(out, histories[0]->alignment[s].data(), 4 * histories[0]->alignment[s].size());

    // 3. Blocking::translate (Frontend.cc:91) with small batches; once more with Options::alignment (Response.alignments)
    slimt::Config config;
    config.max_words = 96;
    slimt::Blocking blocking(config);
    slimt::WordsResponse plain = blocking.translate(model, sources);
    put_sentences(out, plain.target);
    slimt::Options with_alignment;
    with_alignment.alignment = true;
    slimt::WordsResponse aligned = blocking.translate(model, sources, with_alignment);
    if (aligned.target != plain.target || aligned.alignments.size() != sources.size()) throw std::runtime_error("alignment run differs");
    for (size_t i = 0; i < sources.size(); i++) {
      uint32_t rows = aligned.alignments[i].size();
      put(out, &rows, 4);
      for (const auto &d : aligned.alignments[i]) {
        if (d.size() != sources[i].size()) throw std::runtime_error("alignment row width");
        put(out, d.data(), 4 * d.size());
      }
    }

    // 4. Async (Frontend.cc:207-323) with two workers over a model that holds TWO replicas on device 0: concurrent
    // requests, each dealt batch by batch to both replicas; then Async::pivot (Frontend.cc:259-314)
    {
      auto two = std::make_shared<slimt::Model>(slimt::preset::tiny(), paths, std::vector<int>{0, 0});
      config.workers = 2;
      slimt::Config uncached = config;  // (a cached answer comes from whatever batch first produced it; section 5 tests the cache)
      uncached.cache_size = 0;
      slimt::Async async(uncached);
      std::future<slimt::WordsResponse> f1 = async.translate(two, sources);
      // (different options: requests of the same kind that wait together are pooled into one service call -- AggregateBatcher
      // semantics -- and a pooled batch has a different shortlist union; this test compares each request served alone)
      std::future<slimt::WordsResponse> f2 = async.translate(two, slimt::Sentences(sources.begin(), sources.begin() + 1), with_alignment);
      std::future<slimt::WordsResponse> f3 = async.pivot(two, model, sources, with_alignment);
      std::future<slimt::WordsResponse> f4 = async.translate(model, sources);  // a second model on the same device, concurrently
      put_sentences(out, f1.get().target);
      put_sentences(out, f2.get().target);
      slimt::WordsResponse pivoted = f3.get();
      put_sentences(out, pivoted.target);
      if (pivoted.alignments.size() != sources.size()) throw std::runtime_error("pivot alignments missing");
      for (size_t i = 0; i < sources.size(); i++) {
        if (pivoted.alignments[i].size() != pivoted.target[i].size()) throw std::runtime_error("pivot alignment rows");
        for (const auto &d : pivoted.alignments[i])
          if (d.size() != sources[i].size()) throw std::runtime_error("pivot alignment width");
      }
      if (f4.get().target != plain.target) throw std::runtime_error("concurrent request on a shared device context differs");
    }
    // 5. wrap_length and the translation cache (TextProcessor.cc:123-157, Cache.hh, Request.cc:58-78, 120-125)
    {
      slimt::Config wrapped = config;
      wrapped.workers = 1;
      wrapped.wrap_length = 6;
      slimt::Blocking service(wrapped);
      slimt::WordsResponse first = service.translate(model, sources, with_alignment);
      if (first.sentence_begin.size() != sources.size() + 1 || first.sentence_begin.back() != first.source.size())
        throw std::runtime_error("wrap: sentence map");
      if (first.source.size() <= sources.size()) throw std::runtime_error("wrap: nothing was wrapped");
      for (size_t i = 0; i < sources.size(); i++) {
        slimt::Words joined;
        for (size_t k = first.sentence_begin[i]; k < first.sentence_begin[i + 1]; k++) {
          const slimt::Words &seg = first.source[k];
          if (seg.size() > wrapped.wrap_length || seg.back() != 0) throw std::runtime_error("wrap: segment shape");
          const bool last = k + 1 == first.sentence_begin[i + 1];
          joined.insert(joined.end(), seg.begin(), seg.end() - ((last && sources[i].back() == 0) || !last ? 1 : 0));
          if (last && sources[i].back() == 0) joined.push_back(0);
        }
        if (sources[i].size() > wrapped.wrap_length && joined != sources[i]) throw std::runtime_error("wrap: words lost");
      }
      if (service.cache_hits() != 0) throw std::runtime_error("cache: hit on first sight");
      // the same request again: every segment comes from the cache, identical histories
      slimt::WordsResponse again = service.translate(model, sources, with_alignment);
      if (service.cache_hits() != first.source.size()) throw std::runtime_error("cache: expected every segment to hit");
      if (again.target != first.target || again.alignments != first.alignments) throw std::runtime_error("cache: different answer");
      // the segments translated directly by a service without a cache (same batches as the first pass)
      slimt::Config plain_config = wrapped;
      plain_config.cache_size = 0;
      slimt::Blocking no_cache(plain_config);
      slimt::WordsResponse direct = no_cache.translate(model, first.source, with_alignment);
      if (direct.target != first.target || direct.alignments != first.alignments) throw std::runtime_error("wrap: segments differ");
      if (no_cache.cache_hits() != 0) throw std::runtime_error("cache: disabled cache hit");
      // another model does not see this model's records
      auto other = std::make_shared<slimt::Model>(slimt::preset::tiny(), paths);
      if (other->id() == model->id()) throw std::runtime_error("model ids");
      const size_t before = service.cache_hits();
      slimt::WordsResponse other_response = service.translate(other, sources, with_alignment);
      if (service.cache_hits() != before || other_response.target != first.target) throw std::runtime_error("cache: model id not in the key");
    }
    std::printf("host_api_test ok\n");
    return 0;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "host_api_test: %s\n", e.what());
    return 1;
  }
}
