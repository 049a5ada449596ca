// CPU-only checks of the host pieces of include/slimt_b200.hh that need no device: TextProcessor::wrap on word ids
// (slimt/TextProcessor.cc:123-157), cache_key / hash_combine (Request.cc:20-26, Utils.hh:47-57) and AtomicCache
// (Cache.hh:11-58).  Prints the values tests/test_cpp_host_api.py compares with its own restatement.
#include <cstdio>
#include <cstdlib>

#include "slimt_b200.hh"

int main(int argc, char **argv) {
  // wrap: `length` words 1..length followed by EOS (0), wrap_length from argv
  if (argc < 3) return 2;
  const size_t length = std::strtoul(argv[1], nullptr, 10), wrap_length = std::strtoul(argv[2], nullptr, 10);
  slimt::Words sentence;
  for (size_t i = 1; i <= length; i++) sentence.push_back(static_cast<slimt::Word>(i));
  sentence.push_back(0);
  slimt::Sentences segments;
  slimt::wrap(sentence, wrap_length, 0, segments);
  std::printf("segments %zu\n", segments.size());
  for (const auto &s : segments) {
    std::printf("seg");
    for (auto w : s) std::printf(" %u", w);
    std::printf("\n");
  }
  std::printf("key %zu\n", slimt::cache_key(3, sentence));
  std::printf("key0 %zu\n", slimt::cache_key(0, slimt::Words{}));

  // AtomicCache: direct-mapped, a colliding key replaces the record (no probing)
  slimt::AtomicCache<size_t, int> cache(8, 4);
  cache.store(5, 50);
  cache.store(6, 60);
  int ok = cache.find(5).first && cache.find(5).second == 50 && cache.find(6).second == 60 && !cache.find(7).first;
  cache.store(13, 130);  // 13 % 8 == 5: evicts key 5
  ok = ok && !cache.find(5).first && cache.find(13).first && cache.find(13).second == 130 && cache.find(6).first;
  std::printf("cache %s\n", ok ? "ok" : "BAD");
  std::printf("make_cache %d %d\n", slimt::make_cache(0) == nullptr, slimt::make_cache(16) != nullptr);

  // optional: ShortlistGenerator / Shortlist (Shortlist.hh) and the batch plan on a shortlist file -- argv[3] = lex.s2t.bin,
  // argv[4] = vocabulary size; the source words are 1..length
  if (argc >= 5) {
    slimt::io::MmapFile file(argv[3]);
    const size_t vocab = std::strtoul(argv[4], nullptr, 10);
    slimt::ShortlistGenerator generator(slimt::View{file.data(), file.size()}, vocab, /*check=*/true);
    slimt::Shortlist shortlist = generator.generate(sentence);
    std::printf("shortlist");
    for (auto w : shortlist.words()) std::printf(" %u", w);
    std::printf("\n");
    const slimt::Word probe = shortlist.words()[shortlist.words().size() / 2];
    std::printf("maps %d %d %u\n", shortlist.try_forward_map(probe), shortlist.try_forward_map(static_cast<slimt::Word>(vocab + 5)),
                shortlist.reverse_map(static_cast<int>(shortlist.words().size() / 2)));
    slimt::Sentences mixed;
    for (size_t i = 0; i < 10; i++) mixed.push_back(slimt::Words((i * 7) % 11 + 1, 1));
    for (const slimt::BatchPlan &b : slimt::plan_batches(mixed, 24)) {
      std::printf("batch %zu:", b.width);
      for (size_t id : b.ids) std::printf(" %zu", id);
      std::printf("\n");
    }
    bool refused = false;
    try {
      slimt::ShortlistGenerator bad(slimt::View{file.data(), file.size() - 8}, vocab, /*check=*/true);
    } catch (const std::runtime_error &) {
      refused = true;
    }
    std::printf("truncated %s\n", refused ? "refused" : "accepted");
  }
  return ok ? 0 : 1;
}
