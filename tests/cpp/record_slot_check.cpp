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
  if (failures == 0) printf("record_slot_check ok\n");
  return failures != 0;
}
