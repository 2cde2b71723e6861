// record_slot.h -- layout and integrity word of the 64-byte record slots that kernels publish to
// the host (pinned, mapped memory) and to peer GPUs (exchange buffers); shared by device and host
// code.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AMM_HD __host__ __device__ __forceinline__
#else
#define AMM_HD inline
#endif

namespace amm {

constexpr int XCHG_SLOT_WORDS = 16;  // 128 B per (parity, rank) slot: 8 record words, checksum, tag
constexpr int XCHG_MAX_WORLD = 16;   // ranks per exchange; 2 parity banks per buffer
constexpr uint64_t AMM_REC_PEER_TIMEOUT = 4ull;  // record.flags bit: a peer never showed up

// A slot is accepted by its reader only when the tag (seq / xseq) is the expected one AND this
// word matches, so the writer needs no fence between the eight record words: a torn slot is
// simply polled again.  The mix is position-sensitive (multiply-rotate per word, seeded): a slot
// that holds 16-byte pairs of two different records does not pass by accident the way a plain
// XOR would when the stale words happen to cancel (small integer keys and indices do that).
AMM_HD uint64_t mix_checksum(const uint64_t* w /* [8] */, uint64_t seed) {
  uint64_t c = seed;
  for (int i = 0; i < 8; ++i) {
    c = (c ^ w[i]) * 0x9E3779B97F4A7C15ull;
    c = (c << 23) | (c >> 41);
    c += (uint64_t)(i + 1);
  }
  return c;
}
constexpr uint64_t HOST_SLOT_SEED = 0xA5A5C3C3F00DBEEFull;
constexpr uint64_t PEER_SLOT_SEED = 0x5EEDFACE0DDBA11ull;

// Integrity word of a per-CTA grid slot (scan_kernel.cuh: grid_reduce): five data words, seeded
// with the call's sequence number and the slot index, so a slot that still holds an older call's
// record -- or another slot's -- does not pass.  Never 0: a zeroed slot (fresh workspace) cannot
// pass whatever the seed.
constexpr uint64_t GRID_SLOT_SEED = 0xC0FFEE0DD5107ull;
AMM_HD uint64_t grid_slot_check(const uint64_t (&w)[5], uint64_t seq, uint32_t slot) {
  uint64_t c = GRID_SLOT_SEED ^ seq ^ ((uint64_t)slot << 48);
  for (int i = 0; i < 5; ++i) {
    c = (c ^ w[i]) * 0x9E3779B97F4A7C15ull;
    c = (c << 23) | (c >> 41);
    c += (uint64_t)(i + 1);
  }
  return c | 1ull;
}

}  // namespace amm
