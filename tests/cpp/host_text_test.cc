// Text in, text out through the C++ host layer the way slimt's own callers use the reference's services:
// Model from a Package of paths (model, sentencepiece vocabulary, shortlist), Blocking::translate / pivot on strings,
// Async::translate on a string (Frontend.hh:41-78).
// Usage: host_text_test <model.bin> <vocab.spm> <shortlist.bin|-> <split mode> <wrap_length> <max_words> <hex text> ...
// Prints, per source text: "src <annotation>", "seg <segments>", "tgt <hex text> <annotation>", "align <floats>"; then
// the same for the pivot (source -> pivot -> target through the same model) and the Async answer of the first text.
#include <cstdio>
#include <iostream>
#include <sstream>

#include "slimt_b200.hh"

using namespace slimt;  // NOLINT

static std::string unhex(const std::string &h) {
  if (h == "-") return "";
  std::string out;
  for (size_t i = 0; i + 1 < h.size(); i += 2) out.push_back(static_cast<char>(std::stoi(h.substr(i, 2), nullptr, 16)));
  return out;
}
static std::string hex(std::string_view s) {
  if (s.empty()) return "-";
  static const char *d = "0123456789abcdef";
  std::string out;
  for (unsigned char c : s) out.push_back(d[c >> 4]), out.push_back(d[c & 15]);
  return out;
}
static std::string dump(const AnnotatedText &t) {
  std::ostringstream o;
  o << "n=" << t.sentence_count();
  for (size_t s = 0; s < t.sentence_count(); s++) {
    const Range r = t.sentence_as_range(s);
    o << " G" << hex(t.gap(s)) << " S" << r.begin << ":" << r.end << " W";
    for (size_t w = 0; w < t.word_count(s); w++) {
      const Range x = t.word_as_range(s, w);
      o << (w ? "," : "") << x.begin << ":" << x.end;
    }
  }
  o << " G" << hex(t.gap(t.sentence_count()));
  return o.str();
}
static void print(const char *tag, const Response &r) {
  std::cout << tag << " src " << dump(r.source) << "\n" << tag << " tgt " << hex(r.target.text) << " " << dump(r.target) << "\n" << tag << " align";
  char buf[48];
  for (const Alignment &a : r.alignments) {
    std::cout << " [";
    for (const Distribution &row : a)
      for (float p : row) std::snprintf(buf, sizeof(buf), " %a", static_cast<double>(p)), std::cout << buf;
    std::cout << " ]";
  }
  std::cout << "\n";
}

int main(int argc, char **argv) {
  if (argc < 8) return 2;
  try {
    Model::Config mc = preset::tiny();
    mc.split_mode = argv[4];
    Package<std::string> package{argv[1], argv[2], std::string(argv[3]) == "-" ? "" : argv[3], ""};
    auto model = std::make_shared<Model>(mc, package);
    std::cout << "vocab size=" << model->vocabulary().size() << " eos=" << model->config().eos_id << " pad=" << model->config().pad_id << "\n";
    Config config;
    config.wrap_length = std::stoul(argv[5]);
    config.max_words = std::stoul(argv[6]);
    config.workers = 3;
    config.cache_size = 0;
    std::vector<std::string> sources;
    for (int i = 7; i < argc; i++) sources.push_back(unhex(argv[i]));
    Options options;
    options.alignment = true;

    Blocking blocking(config);
    std::vector<Response> responses = blocking.translate(model, sources, options);
    for (const Response &r : responses) print("blocking", r);
    // the segments the service batched (TextProcessor::process once more: deterministic)
    for (const std::string &s : sources) {
      auto [annotated, segments] = model->processor().process(std::string(s), config.wrap_length);
      std::cout << "seg";
      for (const Segment &g : segments) {
        std::cout << " ";
        for (size_t k = 0; k < g.size(); k++) std::cout << (k ? "," : "") << g[k];
      }
      std::cout << "\n";
    }
    // pivot through the same model twice (Frontend.cc:147-205)
    std::vector<Response> pivoted = blocking.pivot(model, model, sources, options);
    for (const Response &r : pivoted) print("pivot", r);
    // Async on the first text, with a cache shared by its workers: a second request is answered from it
    Config with_cache = config;
    with_cache.cache_size = 1024;
    Async async(with_cache);
    Handle h1 = async.translate(model, sources[0], options);
    Response a1 = h1.future().get();
    Handle h2 = async.translate(model, sources[0], options);
    Response a2 = h2.future().get();
    print("async", a1);
    print("cached", a2);
    Handle h3 = async.pivot(model, model, sources[0], options);
    const Handle::Info before = h3.info();
    Response pivoted_async = h3.future().get();
    print("apivot", pivoted_async);
    std::cout << "info parts " << before.parts.q << "\n";
    // a burst of single-line requests: the requests that queue up while the first is served share one service call
    // (AggregateBatcher, Batcher.hh:128-200); each answer is compared with the same text served alone
    {
      std::vector<std::string> burst;
      for (const std::string &s : sources) {
        std::stringstream all(s);
        std::string line;
        while (std::getline(all, line, '\n'))
          if (!line.empty()) burst.push_back(line);
      }
      const size_t base = burst.size();
      for (size_t k = 0; burst.size() < 24; k++) burst.push_back(burst[k % base] + " " + burst[(k + 1) % base]);
      Config no_cache = config;
      no_cache.workers = 1;
      Async pooled(no_cache);
      std::vector<Handle> handles;
      for (const std::string &b : burst) handles.push_back(pooled.translate(model, b, options));
      size_t equal = 0;
      Blocking alone(no_cache);
      for (size_t k = 0; k < burst.size(); k++) {
        Response got = handles[k].future().get();
        Response want = std::move(alone.translate(model, std::vector<std::string>(1, burst[k]), options)[0]);
        // a sentence that never produces EOS is cut at limit_factor x the BATCH's longest sentence (Model.cc:139-143), so
        // alone it may be cut earlier than in company: the shorter answer must be the beginning of the longer
        const std::string &a = got.target.text, &b = want.target.text;
        const bool same_text = a.size() <= b.size() ? b.compare(0, a.size(), a) == 0 : a.compare(0, b.size(), b) == 0;
        bool same_rows = got.alignments.size() == want.alignments.size();
        for (size_t q = 0; same_rows && q < got.alignments.size(); q++) {
          const size_t rows = std::min(got.alignments[q].size(), want.alignments[q].size());
          for (size_t t = 0; t < rows; t++) same_rows = same_rows && got.alignments[q][t] == want.alignments[q][t];
        }
        equal += same_text && same_rows;
      }
      std::cout << "burst requests=" << pooled.requests_served() << " calls=" << pooled.service_calls() << " equal=" << equal << "\n";
    }
    bool refused = false;
    try {
      Options html;
      html.html = true;
      blocking.translate(model, sources, html);
    } catch (const std::runtime_error &) {
      refused = true;
    }
    std::cout << "html " << (refused ? "refused" : "accepted") << "\n";
  } catch (const std::exception &e) {
    std::cerr << "host_text_test: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
