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

int ShortlistGenerator::load(const void* data, size_t bytes) {
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
  const uint64_t expected = sizeof(h) + h.word_to_offset_size * 8 + h.shortlist_size * 4;
  if (expected != bytes) {
    set_error("Shortlist header claims file size should be " + std::to_string(expected) + " but file is " +
              std::to_string(bytes));
    return 1;
  }
  frequent = h.frequent, best = h.best;
  word_to_offset_size = h.word_to_offset_size, shortlist_size = h.shortlist_size;
  const char* p = static_cast<const char*>(data) + sizeof(h);
  word_to_offset = reinterpret_cast<const uint64_t*>(p);
  shortlist = reinterpret_cast<const uint32_t*>(p + word_to_offset_size * 8);
  return 0;
}

std::vector<uint32_t> ShortlistGenerator::generate(const uint32_t* words, size_t n, size_t vocab) const {
  // byte tables instead of the reference's std::vector<bool>: same marks, no bit twiddling on the hot loop
  std::vector<uint8_t> source_table(word_to_offset_size, 0), target_table(vocab, 0);
  size_t ones = 0;
  for (uint32_t i = 0; i < frequent && i < vocab; ++i) target_table[i] = 1, ones++;
  // a large batch's union saturates the vocabulary early: once every id is marked nothing can change
  for (size_t t = 0; t < n && ones < vocab; t++) {
    const uint32_t word = words[t];
    if (word + 1 >= word_to_offset_size || source_table[word]) continue;
    source_table[word] = 1;
    const uint32_t* p = shortlist + word_to_offset[word];
    const uint32_t* e = shortlist + word_to_offset[word + 1];
    for (; p != e; ++p) {
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
  std::vector<uint32_t> indices(ones);
  size_t k = 0;
  for (uint32_t i = 0; i < vocab; i++)
    if (target_table[i]) indices[k++] = i;
  return indices;
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
