// See service.cuh.
#include "service.cuh"

#include <string.h>

#include <algorithm>
#include <string>

#include "engine.cuh"

namespace sb {

namespace {
constexpr uint64_t kShortlistMagic = 0xF11A48D5013417F5ull;
struct ShortlistHeader {
  uint64_t magic, checksum, frequent, best, word_to_offset_size, shortlist_size;
};
}  // namespace

// std::hash<uint64_t> is the identity in libstdc++ (the library the reference is built with), so hash_combine
// (slimt/Utils.hh:46-57) reduces to the boost formula on the raw words.
static uint64_t hash_words(const uint64_t* data, size_t n) {
  uint64_t seed = 0;
  for (size_t i = 0; i < n; i++) seed ^= data[i] + 0x9e3779b9ull + (seed << 6) + (seed >> 2);
  return seed;
}

int ShortlistGenerator::load(const void* data, size_t bytes, size_t vocab, bool check) {
  if (bytes < sizeof(ShortlistHeader)) {
    set_error("Shortlist length too short to have a header: " + std::to_string(bytes));
    return 1;
  }
  ShortlistHeader h;
  memcpy(&h, data, sizeof(h));
  if (h.magic != kShortlistMagic) {
    set_error("Incorrect magic in binary shortlist");
    return 1;
  }
  // header-implied size (Shortlist.cc:58-65); the counts come from the file, so the products must not wrap
  const uint64_t room = bytes - sizeof(h);
  if (h.word_to_offset_size > room / 8 || h.shortlist_size > (room - h.word_to_offset_size * 8) / 4 ||
      sizeof(h) + h.word_to_offset_size * 8 + h.shortlist_size * 4 != bytes) {
    set_error("Shortlist header claims file size should be " + std::to_string(sizeof(h)) + " + 8 * " +
              std::to_string(h.word_to_offset_size) + " + 4 * " + std::to_string(h.shortlist_size) + " but file is " +
              std::to_string(bytes));
    return 1;
  }
  frequent = h.frequent, best = h.best;
  word_to_offset_size = h.word_to_offset_size, shortlist_size = h.shortlist_size;
  const char* p = static_cast<const char*>(data) + sizeof(h);
  word_to_offset = reinterpret_cast<const uint64_t*>(p);
  shortlist = reinterpret_cast<const uint32_t*>(p + word_to_offset_size * 8);
  if (check) {
    // check = true of the reference's loader (Shortlist.cc:68-79, 30-37): checksum over everything after the
    // header's first two words, then content_check()
    const uint64_t* words = static_cast<const uint64_t*>(data) + 2;
    if (hash_words(words, (bytes - 16) / 8) != h.checksum) {
      set_error("checksum check failed: this binary shortlist is corrupted");
      return 1;
    }
    if (word_to_offset_size == 0) {
      set_error("Error: word_to_offset != shortlist_size");
      return 1;
    }
    for (uint64_t i = 0; i + 1 < word_to_offset_size; i++)
      if (word_to_offset[i] >= shortlist_size) {
        set_error("Error: offset table not within shortlist size.");
        return 1;
      }
    if (word_to_offset[word_to_offset_size - 1] != shortlist_size) {
      set_error("Error: word_to_offset != shortlist_size");
      return 1;
    }
    for (uint64_t j = 0; j < shortlist_size; j++)
      if (shortlist[j] >= vocab) {
        set_error("Error: shortlist indices are out of bounds");
        return 1;
      }
  }
  return 0;
}

int ShortlistGenerator::generate(const uint32_t* words, size_t n, size_t vocab, std::vector<uint32_t>* out) const {
  // byte tables instead of the reference's std::vector<bool>: same marks, no bit twiddling on the hot loop.  Unlike the
  // reference with its default check = false, nothing read from the image is trusted: an offset or a target id
  // outside its table is an error, not an out-of-bounds access.
  std::vector<uint8_t> source_table(word_to_offset_size, 0), target_table(vocab, 0);
  size_t ones = 0;
  for (uint32_t i = 0; i < frequent && i < vocab; ++i) target_table[i] = 1, ones++;
  // a large batch's union saturates the vocabulary early: once every id is marked nothing can change
  for (size_t t = 0; t < n && ones < vocab; t++) {
    const uint64_t word = words[t];
    if (shared_vocabulary && word < vocab) {  // `shared_` (Shortlist.cc:132-134): a source word is its own candidate
      ones += 1 - target_table[word];
      target_table[word] = 1;
    }
    if (word + 1 >= word_to_offset_size || source_table[word]) continue;
    source_table[word] = 1;
    const uint64_t begin = word_to_offset[word], end = word_to_offset[word + 1];
    if (begin > end || end > shortlist_size) {
      set_error("Error: offset table not within shortlist size.");
      return 1;
    }
    for (const uint32_t *p = shortlist + begin, *e = shortlist + end; p != e; ++p) {
      if (*p >= vocab) {
        set_error("Error: shortlist indices are out of bounds");
        return 1;
      }
      ones += 1 - target_table[*p];
      target_table[*p] = 1;
    }
  }
  // pad to a multiple of eight with the next unused ids >= frequent (Shortlist.cc:158-164)
  for (size_t i = frequent; i < vocab && ones % 8 != 0; i++) {
    if (!target_table[i]) {
      target_table[i] = 1;
      ones++;
    }
  }
  out->resize(ones);
  size_t k = 0;
  for (uint32_t i = 0; i < vocab; i++)
    if (target_table[i]) (*out)[k++] = i;
  return 0;
}

void Batcher::enqueue(size_t sentence, size_t length) {
  if (length >= bucket_.size()) {
    bucket_.resize(length + 1);
    head_.resize(length + 1, 0);
  }
  bucket_[length].push_back(sentence);
  running_max_ = std::max(running_max_, length);
}

std::vector<size_t> Batcher::generate(size_t* max_length) {
  std::vector<size_t> batch;
  size_t width = 0;
  for (size_t length = 0; length <= running_max_ && length < bucket_.size(); length++) {
    while (head_[length] < bucket_[length].size()) {
      const size_t padded = (batch.size() + 1) * length;
      if (padded <= max_words_ || batch.empty()) {
        batch.push_back(bucket_[length][head_[length]++]);
        width = std::max(width, length);
      } else {
        *max_length = width;
        return batch;
      }
    }
  }
  *max_length = width;
  return batch;
}

}  // namespace sb
