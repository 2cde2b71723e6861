// CPU check of the record-slot integrity word (argminmax_b200/csrc/record_slot.h), plain g++.
// the old plain-XOR checksum accepted a torn slot whenever the stale 16-byte pair
// XORed to the same value as the new one, e.g. old (key = 1, idx = 0) against new (key = 0,
// idx = 1).  The mix must tell every such mixture of two records apart from both originals.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../argminmax_b200/csrc/record_slot.h"

using amm::mix_checksum;

static int failures = 0;
#define CHECK(cond)                                              \
  do {                                                           \
    if (!(cond)) {                                               \
      printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);   \
      ++failures;                                                \
    }                                                            \
  } while (0)

int main() {
  // a record is 8 words: min_key, min_idx, max_key, max_idx, nan_idx, flags, n, seq
  uint64_t oldr[8] = {1, 0, 7, 3, ~0ull, 1, 100, 41};
  uint64_t newr[8] = {0, 1, 7, 3, ~0ull, 1, 100, 42};
  const uint64_t seed = amm::HOST_SLOT_SEED;
  const uint64_t c_new = mix_checksum(newr, seed);
  CHECK(mix_checksum(oldr, seed) != c_new);
  // every torn mixture (16-byte pairs arrive independently): pairs taken from old or new
  int torn_accepted = 0;
  for (int mask = 1; mask < 15; ++mask) {  // at least one pair old, at least one pair new
    uint64_t t[8];
    for (int pair = 0; pair < 4; ++pair) {
      const uint64_t* src = (mask >> pair) & 1 ? oldr : newr;
      t[2 * pair] = src[2 * pair];
      t[2 * pair + 1] = src[2 * pair + 1];
    }
    if (memcmp(t, newr, sizeof(t)) == 0) continue;  // pairs that do not differ
    if (mix_checksum(t, seed) == c_new) ++torn_accepted;
  }
  CHECK(torn_accepted == 0);
  // the XOR of the words is the same for (1,0) and (0,1): the mix must not be
  uint64_t a[8] = {1, 0, 0, 0, 0, 0, 0, 5}, b[8] = {0, 1, 0, 0, 0, 0, 0, 5};
  CHECK(mix_checksum(a, seed) != mix_checksum(b, seed));
  // position sensitivity over all word swaps of a record with distinct words
  uint64_t r[8] = {11, 22, 33, 44, 55, 66, 77, 88};
  const uint64_t c0 = mix_checksum(r, seed);
  for (int i = 0; i < 8; ++i)
    for (int j = i + 1; j < 8; ++j) {
      uint64_t s[8];
      memcpy(s, r, sizeof(s));
      s[i] = r[j];
      s[j] = r[i];
      CHECK(mix_checksum(s, seed) != c0);
    }
  // the seed (the call's tag for peer slots) is part of it
  CHECK(mix_checksum(r, amm::PEER_SLOT_SEED ^ 1) != mix_checksum(r, amm::PEER_SLOT_SEED ^ 2));
  // ---- per-CTA grid slots (grid_reduce): five words + integrity word seeded with (seq, slot)
  {
    using amm::grid_slot_check;
    const uint64_t zero[5] = {0, 0, 0, 0, 0};
    uint64_t d[5] = {(uint64_t)-5, 9, 1234, 77, ~0ull};
    for (uint64_t seq = 1; seq < 2000; ++seq)
      for (uint32_t slot = 0; slot < 600; slot += 37) {
        CHECK(grid_slot_check(zero, seq, slot) != 0);            // a zeroed slot never passes
        CHECK(grid_slot_check(d, seq, slot) != grid_slot_check(d, seq + 1, slot));   // older call
        CHECK(grid_slot_check(d, seq, slot) != grid_slot_check(d, seq, slot + 1));   // other slot
      }
    // torn between two calls' records of the same slot (16-byte pieces: words 01 | 23 | 4c)
    uint64_t o[5] = {3, 9, 10, 20, ~0ull}, n[5] = {9, 3, 20, 10, ~0ull};
    const uint64_t cn = grid_slot_check(n, 8, 5);
    for (int mask = 1; mask < 3; ++mask) {
      uint64_t m[5];
      for (int i = 0; i < 5; ++i) m[i] = ((mask >> (i / 2)) & 1) ? o[i] : n[i];
      CHECK(grid_slot_check(m, 8, 5) != cn);
    }
    // the new data with the OLD call's integrity word
    CHECK(grid_slot_check(n, 7, 5) != cn);
  }
  if (failures == 0) printf("record_slot_check ok\n");
  return failures != 0;
}
