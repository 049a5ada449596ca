// text_bench.cc -- text in, text out through the C++ services on one or more GPUs, with the host's share made visible.
// Usage: text_bench <model.bin> <vocab.spm> <shortlist.bin|-> <text file> <workers> <reps> <max_words> [devices]
// One "step" = Blocking::translate of the file's text (cut into `workers * 4` sources on line boundaries).  Reported:
//   tokenize_s   TextProcessor::process of every source, on `workers` host threads (no GPU)
//   words_s      the same segments through the word-id service: Batcher, shortlist, H2D, GPU, D2H (no text work)
//   e2e_s        the text-level call: tokenize + words + detokenise / Response building
// all as the best of `reps` runs after one warm-up; one JSON line on stdout.
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>

#include "slimt_b200.hh"

using namespace slimt;  // NOLINT
static double seconds(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
  return std::chrono::duration<double>(b - a).count();
}

int main(int argc, char **argv) {
  if (argc < 8) return 2;
  try {
    const std::string sl = std::string(argv[3]) == "-" ? "" : argv[3];
    const size_t workers = std::stoul(argv[5]), reps = std::stoul(argv[6]);
    std::vector<int> devices;
    if (argc > 8) {
      std::stringstream d(argv[8]);
      std::string one;
      while (std::getline(d, one, ',')) devices.push_back(std::stoi(one));
    } else {
      devices.push_back(0);
    }
    std::ifstream f(argv[4], std::ios::binary);
    const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto model = std::make_shared<Model>(preset::tiny(), Package<std::string>{argv[1], argv[2], sl, ""}, devices);
    Config config;
    config.workers = workers;
    config.cache_size = 0;
    config.wrap_length = 128;
    config.max_words = std::stoul(argv[7]);
    // sources: the text cut on line boundaries into workers * 4 pieces
    std::vector<std::string> sources;
    const size_t want = std::max<size_t>(1, workers * 4), chunk = text.size() / want + 1;
    for (size_t at = 0; at < text.size();) {
      size_t end = std::min(text.size(), at + chunk);
      while (end < text.size() && text[end - 1] != '\n') end++;
      sources.push_back(text.substr(at, end - at));
      at = end;
    }
    Blocking service(config);
    double best_tok = 1e30, best_words = 1e30, best_e2e = 1e30;
    size_t sentences = 0, source_tokens = 0, target_tokens = 0, target_bytes = 0;
    for (size_t r = 0; r < reps + 1; r++) {
      // (a) tokenisation alone
      std::vector<Segments> segs(sources.size());
      auto t0 = std::chrono::steady_clock::now();
      {
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        for (size_t t = 0; t < std::max<size_t>(1, workers); t++)
          pool.emplace_back([&]() {
            for (size_t i = next++; i < sources.size(); i = next++)
              segs[i] = std::get<1>(model->processor().process(std::string(sources[i]), config.wrap_length));
          });
        for (auto &t : pool) t.join();
      }
      auto t1 = std::chrono::steady_clock::now();
      // (b) the same segments through the word-id service
      Sentences pool;
      for (const Segments &s : segs) pool.insert(pool.end(), s.begin(), s.end());
      auto t2 = std::chrono::steady_clock::now();
      WordsResponse words = service.translate(model, pool);
      auto t3 = std::chrono::steady_clock::now();
      // (c) the text-level call
      std::vector<Response> responses = service.translate(model, sources, Options());
      auto t4 = std::chrono::steady_clock::now();
      if (r == 0) continue;  // warm-up
      best_tok = std::min(best_tok, seconds(t0, t1)), best_words = std::min(best_words, seconds(t2, t3));
      best_e2e = std::min(best_e2e, seconds(t3, t4));
      sentences = pool.size(), source_tokens = 0, target_tokens = 0, target_bytes = 0;
      for (const Words &w : pool) source_tokens += w.size();
      for (const Words &w : words.target) target_tokens += w.size();
      for (const Response &x : responses) target_bytes += x.target.text.size();
    }
    std::cout << "{\"sentences\": " << sentences << ", \"source_tokens\": " << source_tokens << ", \"target_tokens\": " << target_tokens
              << ", \"source_bytes\": " << text.size() << ", \"target_bytes\": " << target_bytes << ", \"workers\": " << workers
              << ", \"gpus\": " << devices.size() << ", \"tokenize_s\": " << best_tok << ", \"words_s\": " << best_words
              << ", \"e2e_s\": " << best_e2e << "}" << std::endl;
  } catch (const std::exception &e) {
    std::cerr << "text_bench: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
