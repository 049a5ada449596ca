// text_front_driver.cc -- ONE driver, two libraries.  The text front half of the services (Vocabulary, Splitter /
// SentenceStream, TextProcessor, AnnotatedText, Response, combine) is driven by a line protocol on stdin; every
// command answers with one line on stdout.
//
//   * compiled against include/slimt_b200_text.hh  -> tests/cpp/text_front_test   (the product's host code)
//   * compiled with -DSLIMT_TEXT_REFERENCE against the UNMODIFIED reference sources under /root/reference (slimt/
//     Vocabulary.cc, TextProcessor.cc, Splitter.cc, Regex.cc, Annotation.cc, Response.cc, Request.cc + the vendored
//     sentencepiece 0.2.00)                     -> oracle/_ref/text_ref          (test infrastructure; oracle/Makefile)
//
// tests/golden/make_text_golden.py records the reference binary's answers; tests/test_text_front.py replays the
// commands through the product binary and compares the answers byte for byte.
//
//   vocab <path>
//   encode <hex text>                          Vocabulary::encode          -> ids | byte ranges in the text
//   decode <id> ...                            Vocabulary::decode          -> hex text | byte ranges
//   split <sentence|paragraph|wrapped_text> <hex prefixes or -> <hex text> SentenceStream -> sentence ranges ("-" = separator)
//   process <mode> <wrap_length> <hex text>    TextProcessor::process      -> annotation | segments
//   respond <mode> <wrap_length> <hex text> <ids,ids;ids,...>   Request::complete  -> source annotation | target text + annotation
//   pivot <mode> <wrap_length> <hex text> <targets of model 1> <targets of model 2>   process(AnnotatedText&) + combine
//   recode <mode> <wrap_length> <hex text>     AnnotatedText::to(UTF8) and back -> both offset tables
//   bench <mode> <wrap_length> <file> <reps>   timing: TextProcessor::process of the file's text, then Vocabulary::decode of
//                                              every segment (the two host-side halves of a request) -> counts and seconds
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <optional>
#include <sstream>
#include <string>
#include <vector>

#ifdef SLIMT_TEXT_REFERENCE
#include "slimt/Aligned.hh"
#include "slimt/Annotation.hh"
#include "slimt/Request.hh"
#include "slimt/Response.hh"
#include "slimt/Splitter.hh"
#include "slimt/TextProcessor.hh"
#include "slimt/Types.hh"
#include "slimt/Vocabulary.hh"
#else
#include "slimt_b200_text.hh"
#endif

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
static std::string range(const Range &r) { return std::to_string(r.begin) + ":" + std::to_string(r.end); }

// everything the public interface of an AnnotatedText tells about its annotation
static std::string dump(const AnnotatedText &t) {
  std::ostringstream o;
  o << "n=" << t.sentence_count();
  for (size_t s = 0; s < t.sentence_count(); s++) {
    o << " G" << hex(t.gap(s)) << " S" << range(t.sentence_as_range(s)) << " W";
    for (size_t w = 0; w < t.word_count(s); w++) o << (w ? "," : "") << range(t.word_as_range(s, w));
  }
  o << " G" << hex(t.gap(t.sentence_count()));
  return o.str();
}

static std::vector<Words> parse_targets(const std::string &spec) {
  std::vector<Words> out;
  if (spec == "-") return out;
  std::stringstream all(spec);
  std::string one;
  while (std::getline(all, one, ';')) {
    Words w;
    std::stringstream ids(one);
    std::string id;
    while (std::getline(ids, id, ',')) w.push_back(static_cast<Word>(std::stoul(id)));
    out.push_back(std::move(w));
  }
  return out;
}

// a deterministic stand-in for the decoder's attention rows: T rows over S source tokens, each summing to ~1
static Alignment fake_alignment(size_t T, size_t S, size_t salt) {
  Alignment a(T, Distribution(S, 0.0F));
  for (size_t t = 0; t < T; t++) {
    float sum = 0.0F;
    for (size_t s = 0; s < S; s++) a[t][s] = static_cast<float>((t * 7 + s * 13 + salt) % 17 + 1), sum += a[t][s];
    for (size_t s = 0; s < S; s++) a[t][s] /= sum;
  }
  return a;
}

static Histories make_histories(const Segments &segments, const std::vector<Words> &targets, size_t salt) {
  Histories h;
  for (size_t i = 0; i < segments.size(); i++) {
    auto hyp = std::make_shared<Hypothesis>();
    hyp->target = targets.at(i);
    hyp->alignment = fake_alignment(hyp->target.size(), segments[i].size(), salt + i);
    h.push_back(std::move(hyp));
  }
  return h;
}

// the Response of a request whose sentences came back as `histories`
static Response respond(AnnotatedText &&source, Segments &&segments, const Histories &histories, const Vocabulary &vocabulary) {
#ifdef SLIMT_TEXT_REFERENCE
  Response out;
  std::optional<TranslationCache> cache;
  auto request = std::make_shared<Request>(0, 0, std::move(source), std::move(segments), vocabulary, cache,
                                           [&out](Response &&r) {
                                             out = std::move(r);
                                             return nullptr;
                                           });
  for (size_t i = 0; i < histories.size(); i++) request->process(i, histories[i]);
  return out;
#else
  (void)segments;
  return make_response(std::move(source), histories, vocabulary);
#endif
}

static std::unique_ptr<TextProcessor> make_processor(const std::string &mode, const Vocabulary &vocabulary) {
#ifdef SLIMT_TEXT_REFERENCE
  return std::make_unique<TextProcessor>(mode, vocabulary, Aligned());
#else
  return std::make_unique<TextProcessor>(mode, vocabulary);
#endif
}

int main() {
  std::unique_ptr<Vocabulary> vocabulary;
  std::string line;
  while (std::getline(std::cin, line)) {
    std::stringstream in(line);
    std::string cmd;
    in >> cmd;
    std::ostringstream out;
    if (cmd == "vocab") {
      std::string path;
      in >> path;
      vocabulary = std::make_unique<Vocabulary>(path);
      out << "vocab size=" << vocabulary->size() << " eos=" << vocabulary->eos_id() << " pad=" << vocabulary->pad_id();
    } else if (cmd == "encode") {
      std::string h;
      in >> h;
      const std::string text = unhex(h);
      auto [words, views] = vocabulary->encode(text, false);
      out << "ids";
      for (Word w : words) out << " " << w;
      out << " |";
      for (auto v : views) out << " " << (v.data() - text.data()) << ":" << (v.data() - text.data() + v.size());
    } else if (cmd == "decode") {
      Words words;
      Word w;
      while (in >> w) words.push_back(w);
      std::string text;
      Views views = vocabulary->decode(words, text, /*ignore_eos=*/false);
      out << "text " << hex(text) << " |";
      for (auto v : views) out << " " << (v.data() - text.data()) << ":" << (v.data() - text.data() + v.size());
    } else if (cmd == "split") {
      std::string mode, hp, ht;
      in >> mode >> hp >> ht;
      const std::string prefixes = unhex(hp), text = unhex(ht);
      Splitter splitter;
      if (!prefixes.empty()) splitter.load_from_serialized(prefixes);
      using M = SentenceStream::splitmode;
      const M m = mode == "sentence" ? M::OneSentencePerLine : (mode == "paragraph" ? M::OneParagraphPerLine : M::WrappedText);
      SentenceStream stream(std::string_view(text.data(), text.size()), splitter, m);
      std::string_view snt;
      out << "snt";
      while (stream >> snt) {
        if (snt.data() == nullptr) {
          out << " -";
        } else {
          out << " " << (snt.data() - text.data()) << ":" << (snt.data() - text.data() + snt.size());
        }
      }
      if (!stream.error_message().empty()) out << " error";
    } else if (cmd == "process" || cmd == "respond" || cmd == "pivot" || cmd == "recode") {
      std::string mode, ht, spec1, spec2;
      size_t wrap = 0;
      in >> mode >> wrap >> ht >> spec1 >> spec2;
      auto processor = make_processor(mode, *vocabulary);
      auto [annotated, segments] = processor->process(unhex(ht), wrap);
      if (cmd == "process") {
        out << "ann " << dump(annotated) << " | seg";
        for (const Segment &s : segments) {
          out << " ";
          for (size_t i = 0; i < s.size(); i++) out << (i ? "," : "") << s[i];
        }
      } else if (cmd == "recode") {
        out << "bytes " << dump(annotated);
        annotated.to(Encoding::UTF8);
        out << " | utf8";
        for (size_t s = 0; s < annotated.sentence_count(); s++)
          for (size_t w = 0; w < annotated.word_count(s); w++) out << " " << range(annotated.annotation.word(s, w));
        annotated.to(Encoding::Byte);
        out << " | back " << dump(annotated);
      } else {
        Segments copy = segments;
        const Histories h1 = make_histories(copy, parse_targets(spec1), 3);
        Response first = respond(std::move(annotated), std::move(segments), h1, *vocabulary);
        if (cmd == "respond") {
          out << "src " << dump(first.source) << " | tgt " << hex(first.target.text) << " " << dump(first.target) << " | align "
              << first.alignments.size();
        } else {
          auto [pivot_annotated, pivot_segments] = processor->process(first.target);
          out << "piv " << dump(pivot_annotated) << " | seg";
          for (const Segment &s : pivot_segments) {
            out << " ";
            for (size_t i = 0; i < s.size(); i++) out << (i ? "," : "") << s[i];
          }
          Segments copy2 = pivot_segments;
          const Histories h2 = make_histories(copy2, parse_targets(spec2), 11);
          Response second = respond(std::move(pivot_annotated), std::move(pivot_segments), h2, *vocabulary);
          Response combined = combine(std::move(first), std::move(second));
          out << " | tgt " << hex(combined.target.text) << " | align";
          char buf[64];
          for (const Alignment &a : combined.alignments) {
            out << " [";
            for (const Distribution &row : a)
              for (float p : row) std::snprintf(buf, sizeof(buf), " %a", static_cast<double>(p)), out << buf;
            out << " ]";
          }
        }
      }
    } else if (cmd == "bench") {
      std::string mode, path;
      size_t wrap = 0, reps = 1;
      in >> mode >> wrap >> path >> reps;
      std::ifstream f(path, std::ios::binary);
      const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
      auto processor = make_processor(mode, *vocabulary);
      size_t sentences = 0, tokens = 0, decoded_bytes = 0;
      double t_process = 0, t_decode = 0;
      for (size_t r = 0; r < reps; r++) {
        auto t0 = std::chrono::steady_clock::now();
        auto [annotated, segments] = processor->process(std::string(text), wrap);
        auto t1 = std::chrono::steady_clock::now();
        for (const Segment &s : segments) {
          std::string decoded;
          Views views = vocabulary->decode(s, decoded, /*ignore_eos=*/false);
          decoded_bytes += decoded.size() + views.size();
        }
        auto t2 = std::chrono::steady_clock::now();
        t_process += std::chrono::duration<double>(t1 - t0).count(), t_decode += std::chrono::duration<double>(t2 - t1).count();
        sentences += segments.size();
        for (const Segment &s : segments) tokens += s.size();
      }
      out << "bench bytes=" << text.size() * reps << " sentences=" << sentences << " tokens=" << tokens << " process_s=" << t_process
          << " decode_s=" << t_decode << " check=" << decoded_bytes;
#ifndef SLIMT_TEXT_REFERENCE
    } else if (cmd == "loadfuzz") {
      // product only: damaged vocabulary files must end in std::runtime_error (or load), never in a crash
      std::string path;
      uint64_t seed = 1;
      size_t n = 0;
      in >> path >> seed >> n;
      std::ifstream f(path, std::ios::binary);
      const std::string good((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
      size_t loaded = 0, refused = 0;
      auto next = [&seed]() {
        seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
        return seed >> 33;
      };
      for (size_t k = 0; k < n; k++) {
        std::string bad = good;
        const int kind = static_cast<int>(next() % 4);
        if (kind == 0) bad.resize(next() % (bad.size() + 1));                                   // truncated
        if (kind == 1) for (int j = 0; j < 8; j++) bad[next() % bad.size()] = static_cast<char>(next());  // flipped bytes
        if (kind == 2) bad.insert(next() % bad.size(), std::string(next() % 64, static_cast<char>(next())));  // inserted run
        if (kind == 3) for (int j = 0; j < 64; j++) bad[(good.size() - 300000 + next() % 290000) % bad.size()] = static_cast<char>(next());  // charsmap / trie damage
        try {
          Vocabulary v(View{bad.data(), bad.size()});
          auto [w, r] = v.encode("Hello wor\xef\xbc\xa1ld \xef\xac\x81ne. 12", true);
          std::string text;
          v.decode(w, text, false);
          loaded++;
        } catch (const std::runtime_error &) {
          refused++;
        } catch (const std::length_error &) {
          refused++;
        } catch (const std::bad_alloc &) {
          refused++;
        }
      }
      out << "loadfuzz loaded=" << loaded << " refused=" << refused;
#endif
    } else if (cmd.empty()) {
      continue;
    } else {
      out << "unknown command";
    }
    std::cout << out.str() << "\n" << std::flush;
  }
  return 0;
}
