/* TEST INFRASTRUCTURE — not product code.
 *
 * Force-included (g++ -include) when building the UNMODIFIED reference
 * translation unit /root/reference/src/pbsim.cpp into oracle/_ref/pbsim_logrand.
 * It interposes two libc calls by macro, without touching the reference source:
 *
 *   rand()    -> every draw is appended (int32, host endian) to $PBSIM_DRAW_LOG
 *   sprintf() -> when the format is one of the read-id formats
 *                ("%s%ld_%ld" pbsim.cpp:2319, "%s%ld/%ld/%ld" :2322, "%s_%ld" :2951)
 *                the current draw count (int64) is appended to $PBSIM_MARK_LOG.
 *                A mark is therefore written once per emitted (read, pass), after
 *                all of its draws: mark[k-1] is the index of the first draw of
 *                subread k (subread 0 starts at draw 0).
 *
 * Program output is unchanged by the interposition (checked by tests).
 */
#ifndef PBSIM_LOGRANDOM_H
#define PBSIM_LOGRANDOM_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static FILE *pbh_draw_fp = NULL, *pbh_mark_fp = NULL;
static long long pbh_ndraws = 0;
static int pbh_init_done = 0;

static void pbh_fini(void) {
  if (pbh_draw_fp) fclose(pbh_draw_fp);
  if (pbh_mark_fp) fclose(pbh_mark_fp);
}

static void pbh_init(void) {
  const char *p;
  pbh_init_done = 1;
  if ((p = getenv("PBSIM_DRAW_LOG")) != NULL) pbh_draw_fp = fopen(p, "wb");
  if ((p = getenv("PBSIM_MARK_LOG")) != NULL) pbh_mark_fp = fopen(p, "wb");
  atexit(pbh_fini);
}

static inline int pbh_rand(void) {
  int v = (rand)();
  if (!pbh_init_done) pbh_init();
  if (pbh_draw_fp) fwrite(&v, sizeof(int), 1, pbh_draw_fp);
  pbh_ndraws++;
  return v;
}

static inline void pbh_mark(const char *fmt) {
  if (!pbh_init_done) pbh_init();
  if (pbh_mark_fp &&
      (strcmp(fmt, "%s%ld_%ld") == 0 || strcmp(fmt, "%s%ld/%ld/%ld") == 0 ||
       strcmp(fmt, "%s_%ld") == 0 || strcmp(fmt, "%s_%ld/%ld/%ld") == 0 ||
       strcmp(fmt, "%s/%ld/%ld") == 0)) {
    fwrite(&pbh_ndraws, sizeof(long long), 1, pbh_mark_fp);
  }
}

#define rand() pbh_rand()
#define sprintf(buf, fmt, ...) (pbh_mark(fmt), (sprintf)(buf, fmt, ##__VA_ARGS__))
#endif
