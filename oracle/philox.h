/* TEST INFRASTRUCTURE (oracle) — Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11),
 * restated from the paper.  The engine's PHILOX mode (no counterpart in the reference,
 * which only has libc rand()) addresses every draw by
 *   key     = (seed, sequence number)
 *   counter = (position, block | pass << 16, read id, domain)
 * see DESIGN.md "Philox draw addressing".  tests pin this file against the Random123
 * known-answer vectors.
 */
#ifndef ORC_PHILOX_H
#define ORC_PHILOX_H
#include <stdint.h>

static inline void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  int r;
  for (r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif
