/* TEST INFRASTRUCTURE (oracle) — restatement of glibc's srand()/rand().
 *
 * The reference draws every random number with libc rand() after srand(seed)
 * (/root/reference/src/pbsim.cpp:543 and e.g. :2174,2183,2189,2216-2232).
 * The algorithm lives in a third-party dependency that is not in /root/reference:
 * GNU libc 2.39 (stdlib/random_r.c), generator TYPE_3: a degree-31 additive
 * feedback generator  r[i] = r[i-31] + r[i-3]  over int32 with the state seeded by
 * the Lehmer step  r[i] = 16807 * r[i-1] mod (2^31-1)  and 310 discarded outputs;
 * each output is  (uint32)r[i] >> 1.  Restated here from the published algorithm so
 * that golden draw streams can be regenerated from a seed on a box that has neither
 * the reference nor (necessarily) the same libc.  tests/test_oracle_rand.py pins
 * it against the system rand().
 */
#include <stdint.h>
#include "pbsim_oracle.h"

void orc_glibc_srand(orc_glibc_rand_t *g, uint32_t seed) {
  int32_t *r = g->r;
  int64_t word;
  int i;
  if (seed == 0) seed = 1;
  r[0] = (int32_t)seed;
  for (i = 1; i < 31; i++) {
    /* 16807 * r[i-1] % 2147483647 computed without overflow (Schrage) */
    int64_t hi = r[i - 1] / 127773;
    int64_t lo = r[i - 1] % 127773;
    word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    r[i] = (int32_t)word;
  }
  g->f = 3;  /* front = state + SEP_3 */
  g->b = 0;  /* rear  = state */
  for (i = 0; i < 310; i++) (void)orc_glibc_rand(g);
}

int32_t orc_glibc_rand(orc_glibc_rand_t *g) {
  uint32_t v;
  g->r[g->f] = (int32_t)((uint32_t)g->r[g->f] + (uint32_t)g->r[g->b]);
  v = (uint32_t)g->r[g->f] >> 1;
  if (++g->f >= 31) g->f = 0;
  if (++g->b >= 31) g->b = 0;
  return (int32_t)v;
}

void orc_glibc_rand_fill(uint32_t seed, int64_t n, int32_t *out) {
  orc_glibc_rand_t g;
  int64_t i;
  orc_glibc_srand(&g, seed);
  for (i = 0; i < n; i++) out[i] = orc_glibc_rand(&g);
}
