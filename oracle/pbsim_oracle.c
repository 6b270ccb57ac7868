/* TEST INFRASTRUCTURE — CPU oracle for the PBSIM3 read-generation hot path.
 * See pbsim_oracle.h for the role of this file and how its parity is pinned.
 *
 * Every function cites the lines of /root/reference/src/pbsim.cpp ("ref:") whose
 * behaviour it restates.  The restatement is structured around a draw source with
 * purpose-named draws so that the same control flow serves three modes:
 *   glibc   srand(seed)/rand() restated (glibc_rand.c): reproduces the reference
 *   replay  the same, but reading a captured draw log
 *   philox  the engine's counter-addressed Philox4x32-10 stream (no reference
 *           counterpart; statistical parity only) — see DESIGN.md for addressing.
 */
#include "pbsim_oracle.h"
#include "philox.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LINE_MAX_BYTES 10240 /* ref: BUF_SIZE :20 */
#define NACC (ORC_ACC_MAX + 1)
#define NST (ORC_STATE_MAX + 1)

/* ------------------------------------------------------------------ buffers */
typedef struct {
  char *p;
  int64_t n, cap;
} buf_t;

static void buf_reserve(buf_t *b, int64_t extra) {
  if (b->n + extra <= b->cap) return;
  int64_t cap = b->cap ? b->cap : 1 << 16;
  while (cap < b->n + extra) cap *= 2;
  b->p = (char *)realloc(b->p, (size_t)cap);
  b->cap = cap;
}
static void buf_put(buf_t *b, const void *src, int64_t n) {
  buf_reserve(b, n);
  memcpy(b->p + b->n, src, (size_t)n);
  b->n += n;
}
static void buf_puts(buf_t *b, const char *s) { buf_put(b, s, (int64_t)strlen(s)); }
static void buf_pad(buf_t *b, int n) {
  while (n-- > 0) buf_put(b, " ", 1);
}
static void buf_long(buf_t *b, const char *pre, long v, const char *post) {
  char tmp[64];
  snprintf(tmp, sizeof tmp, "%s%ld%s", pre, v, post);
  buf_puts(b, tmp);
}

/* ------------------------------------------------------------------ draw source */
enum { RNG_GLIBC = 0, RNG_REPLAY = 1, RNG_PHILOX = 2 };

typedef struct {
  int mode;
  orc_glibc_rand_t g;
  const int32_t *log;
  int64_t nlog;
  int64_t cur; /* draws consumed (stream modes) */
  int32_t *rec;
  int64_t rec_cap;
  int exhausted;
  /* philox */
  uint32_t key[2];
  uint32_t read_id, pass, pos;
  uint32_t w[4];  /* block 0 of the current position */
  uint32_t cw[4]; /* chain block: state draws of positions 4*cidx .. 4*cidx+3 */
  uint32_t cidx;
  int cvalid;
  /* addressing scheme of the per-position draws: 0 = errhmm (one block per column + chain stream),
   * 1 = qshmm / sample "v2" (DESIGN.md 2.5): error stream = domain 1, one block per TWO positions
   * (block pos>>1, words 2*(pos&1) = X and 2*(pos&1)+1 = Y); quality stream = domain 2, one block per FOUR
   * positions (block pos>>2, word pos&3 = S: high half state draw, low half emission draw) */
  int scheme;
  uint32_t ew[4]; /* error block of positions 2*eidx, 2*eidx+1 */
  uint32_t eidx;
  int evalid;
} rng_t;

static uint32_t stream_next(rng_t *r) {
  int32_t v;
  if (r->mode == RNG_GLIBC) {
    v = orc_glibc_rand(&r->g);
    if (r->cur >= r->rec_cap) {
      r->rec_cap = r->rec_cap ? r->rec_cap * 2 : 1 << 20;
      r->rec = (int32_t *)realloc(r->rec, (size_t)r->rec_cap * sizeof(int32_t));
    }
    r->rec[r->cur] = v;
  } else {
    if (r->cur >= r->nlog) {
      r->exhausted = 1;
      v = 0;
    } else {
      v = r->log[r->cur];
    }
  }
  r->cur++;
  return (uint32_t)v;
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

static void philox_at(const rng_t *r, uint32_t pos, uint32_t blk, uint32_t domain, uint32_t out[4]) {
  uint32_t ctr[4];
  ctr[0] = pos;
  ctr[1] = blk | (r->pass << 16);
  ctr[2] = r->read_id;
  ctr[3] = domain;
  orc_philox4x32_10(ctr, r->key, out);
}

/* planner draws (ref: :2174, :2183, :2189) */
static void d_plan_begin(rng_t *r, uint32_t read_id) {
  r->read_id = read_id;
  r->pass = 0;
  r->cvalid = 0;
  r->evalid = 0;
  if (r->mode == RNG_PHILOX) philox_at(r, 0, 0, 0, r->w);
}
static uint32_t d_plan_len(rng_t *r, uint32_t mod) {
  return r->mode == RNG_PHILOX ? mulhi32(r->w[0], mod) : stream_next(r) % mod;
}
static uint32_t d_plan_acc(rng_t *r, uint32_t mod) {
  return r->mode == RNG_PHILOX ? mulhi32(r->w[1], mod) : stream_next(r) % mod;
}
static uint64_t d_plan_off(rng_t *r, uint64_t span) {
  if (r->mode == RNG_PHILOX) {
    uint64_t u = ((uint64_t)r->w[2] << 32) | r->w[3];
    return (uint64_t)(((unsigned __int128)u * span) >> 64);
  }
  return (uint64_t)stream_next(r) % span;
}

/* per-position draws */
static void d_begin(rng_t *r, uint32_t pass, uint32_t pos) {
  if (r->pass != pass) r->cvalid = r->evalid = 0;
  r->pass = pass;
  r->pos = pos;
  if (r->mode != RNG_PHILOX) return;
  if (r->scheme == 1) {
    if (!r->evalid || r->eidx != (pos >> 1)) {
      philox_at(r, pos >> 1, 0, 1, r->ew);
      r->eidx = pos >> 1;
      r->evalid = 1;
    }
    r->w[0] = r->ew[(pos & 1u) * 2u];      /* X: error draw, 3-way choice in the low 12 bits */
    r->w[1] = r->ew[(pos & 1u) * 2u + 1u]; /* Y: deletion draws, 8-way choice bits 0-2, 4-way choice bits 3-4 */
  } else {
    philox_at(r, pos, 0, 1, r->w);
  }
}
/* quality-stream word S of the current position (scheme 1) */
static uint32_t qs_word(rng_t *r) {
  if (!r->cvalid || r->cidx != (r->pos >> 2)) {
    philox_at(r, r->pos >> 2, 0, 2, r->cw);
    r->cidx = r->pos >> 2;
    r->cvalid = 1;
  }
  return r->cw[r->pos & 3u];
}
/* PHILOX scheme 1: a 32-bit word W "hits" a threshold of t millionths iff W < T32(t),
 * T32(t) = min(ceil(t * 2^32 / 10^6), 2^32 - 1); for t < 10^6 that is mulhi32(W, 10^6) < t exactly. */
static uint32_t t32_of(double thr) {
  double c = ceil(thr);
  uint64_t t, v;
  if (!(c >= 0)) c = 0;
  if (c > 1000000.0) c = 1000000.0;
  t = (uint64_t)c;
  v = (t * 4294967296ull + 999999ull) / 1000000ull;
  return v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
}
/* does draw `d` (stream modes: a value 0..999999; scheme 1: a raw word) fall below `thr` millionths? */
static int d_lt(const rng_t *r, uint32_t d, double thr) {
  if (r->mode == RNG_PHILOX && r->scheme == 1) return d < t32_of(thr);
  return (double)d < thr;
}
/* HMM state draw (init2state / tran2state index) of the current position.  PHILOX mode: the chain has its own
 * stream, domain 2, one block per 4 consecutive positions (block pos>>2, word pos&3), so that the state chain can
 * be advanced without generating the other draws of a position. */
static uint32_t d_state(rng_t *r, uint32_t mod) {
  if (r->mode != RNG_PHILOX) return stream_next(r) % mod;
  if (r->scheme == 1) return mulhi32(qs_word(r) & 0xFFFF0000u, mod); /* high half: (S >> 16) * mod >> 16 */
  if (!r->cvalid || r->cidx != (r->pos >> 2)) {
    philox_at(r, r->pos >> 2, 0, 2, r->cw);
    r->cidx = r->pos >> 2;
    r->cvalid = 1;
  }
  return mulhi32(r->cw[r->pos & 3u], mod);
}
static uint32_t d_w1(rng_t *r, uint32_t mod) {
  return r->mode == RNG_PHILOX ? mulhi32(r->w[1], mod) : stream_next(r) % mod;
}
static uint32_t d_w2(rng_t *r, uint32_t mod) {
  return r->mode == RNG_PHILOX ? mulhi32(r->w[2], mod) : stream_next(r) % mod;
}
static uint32_t d_w3(rng_t *r, uint32_t mod) {
  return r->mode == RNG_PHILOX ? mulhi32(r->w[3], mod) : stream_next(r) % mod;
}
/* qshmm emission / freq2qc index (ref: :2226, :2230) */
static uint32_t d_qs_emis(rng_t *r, uint32_t mod) {
  if (r->mode != RNG_PHILOX) return stream_next(r) % mod;
  return mulhi32(qs_word(r) << 16, mod); /* low half: (S & 0xFFFF) * mod >> 16 */
}
static uint32_t d_qs_freq(rng_t *r, uint32_t mod) {
  return r->mode == RNG_PHILOX ? mulhi32(qs_word(r), mod) : stream_next(r) % mod;
}
/* qshmm / sample error draw (ref: rand() % 1000000, :2234); compare with d_lt */
static uint32_t d_qs_err(rng_t *r) {
  return r->mode == RNG_PHILOX ? r->w[0] : stream_next(r) % 1000000;
}
static uint32_t d_choice3(rng_t *r) {
  return r->mode == RNG_PHILOX ? ((r->w[0] & 0xFFFu) * 3u) >> 12 : stream_next(r) % 3;
}
static uint32_t d_choice4(rng_t *r) {
  if (r->mode != RNG_PHILOX) return stream_next(r) % 4;
  return r->scheme == 1 ? (r->w[1] >> 3) & 3u : (r->w[0] >> 12) & 3u;
}
static uint32_t d_choice8(rng_t *r) {
  return r->mode == RNG_PHILOX ? r->w[1] & 7u : stream_next(r) % 8;
}
static uint32_t d_mag3(rng_t *r) { /* errhmm: rand()%3+1 (ref: :3895) */
  return r->mode == RNG_PHILOX ? (((r->w[3] & 0xFFFu) * 3u) >> 12) + 1 : stream_next(r) % 3 + 1;
}
/* murmur3 finaliser (bijection on 32 bits) */
static uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
/* qshmm deletion draw number j (0-based) after the current position (ref: :2270).
 * PHILOX mode: j = 0 is the position's word Y; the rare later draws are derived from it (scheme 1: a
 * multiply-add chain, errhmm scheme 0: the murmur finaliser; engine definition, DESIGN.md "Philox draw addressing"). */
static uint32_t d_del(rng_t *r, uint32_t j) {
  if (r->mode != RNG_PHILOX) return stream_next(r) % 1000000;
  if (r->scheme == 1) { /* raw word (compare with d_lt): d_0 = Y, d_j = d_{j-1} * 0x9E3779B1 + 0x7F4A7C15 (one multiply-add per further draw) */
    uint32_t d = r->w[1], k;
    for (k = 0; k < j; k++) d = d * 0x9E3779B1u + 0x7F4A7C15u;
    return d;
  }
  if (j == 0) return mulhi32(r->w[3], 1000000);
  return mulhi32(fmix32(r->w[3] + j * 0x9E3779B9u), 1000000);
}

/* ------------------------------------------------------------------ context */
struct orc_ctx {
  char err[256];
  /* parameters */
  int method, pass_num;
  double accuracy_mean;
  long len_min, len_max;
  double len_mean, len_sd;
  long sub_ratio, ins_ratio, del_ratio;
  double sub_rate, ins_rate, del_rate;
  double hp_del_bias_opt;
  char id_prefix[256];
  /* model, laid out flat exactly like the reference's arrays so that out-of-range
   * state numbers alias the same cells (ref: struct qshmm_t/errhmm_t :160-178) */
  double *ip;   /* [NACC][NST] */
  double *ep;   /* [NACC][NST][94] (errhmm: [NACC][NST][4]) */
  double *tp;   /* [NACC][NST][NST] */
  int ep_cols;
  int exist[NACC];
  int state_max[NACC];
  int acc_min, acc_max;
  int model_loaded;
  /* Phred / uniform tables (ref: :546-578) */
  double qc_prob[ORC_NQV];
  double uni_ep[NACC][ORC_NQV];
  long sub_thre[ORC_NQV], ins_thre[ORC_NQV], del_thre[ORC_NQV];
  /* quantised tables, 1-based like the reference */
  long *prob2len;  /* [100001] */
  long *prob2acc;  /* [100001] */
  long len_rand_value, accuracy_rand_value;
  long tab_acc_lo, tab_acc_hi;
  /* qshmm (resolution 100) and freq2qc (1000) */
  uint8_t (*qs_init)[101];         /* [NACC][101] */
  uint8_t (*qs_emis)[NST][101];    /* [NACC][NST][101] */
  uint8_t (*qs_tran)[NST][101];
  uint8_t (*qs_freq)[1001];        /* [NACC][1001] */
  long mod_init[NACC], mod_freq[NACC];
  long mod_emis[NACC][NST], mod_tran[NACC][NST];
  /* errhmm (resolution 1000) */
  uint8_t (*er_init)[1001];        /* [NACC][1001] */
  uint8_t (*er_emis)[NST][1001];
  uint8_t (*er_tran)[NST][1001];
  long er_del[NACC][NST];
  int tables_built;
  /* genome */
  char *seq;
  int16_t *hp; /* hp[-1] is readable: allocated with one leading cell */
  int16_t *hp_alloc;
  int64_t glen;
  int seq_num;
  long hpfreq[12];  /* [11] aliases hp_del_bias[0] in the reference build, see set_sequence */
  double bias[12];
  /* record naming: WGS writes "ref" and ids "<prefix><seqnum>_<read>"; the transcript / template
   * strategies write the sequence's own name and ids "<prefix>_<read>" (ref: :2951, :2968-2986, :3469, :3486) */
  int set_mode;            /* 0 WGS, 1 transcript, 2 template */
  const char *rec_name;
  int rec_name_width;      /* digit_num1[0]: strlen(id) for transcripts, 3 otherwise */
  /* window scratch */
  char *w_seq, *read_seq, *qual, *maf_seq, *maf_ref;
  int16_t *w_hp_alloc, *w_hp;
  int64_t w_cap;
  /* draw source */
  rng_t rng;
  /* outputs */
  buf_t out_reads, out_maf;
  orc_stats_t st;
  int64_t *freq_len;
  int64_t freq_len_n;
  int64_t *freq_acc;
  orc_readinfo_t *info;
  int64_t info_n, info_cap;
};

static int fail(orc_ctx *c, const char *msg) {
  snprintf(c->err, sizeof c->err, "%s", msg);
  return -1;
}

orc_ctx *orc_new(void) {
  orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
  int i, j;
  double prob, rate;
  c->ip = (double *)calloc((size_t)NACC * NST, sizeof(double));
  c->tp = (double *)calloc((size_t)NACC * NST * NST, sizeof(double));
  c->prob2len = (long *)calloc(100001, sizeof(long));
  c->prob2acc = (long *)calloc(100001, sizeof(long));
  c->qs_init = calloc(NACC, sizeof(*c->qs_init));
  c->qs_emis = calloc(NACC, sizeof(*c->qs_emis));
  c->qs_tran = calloc(NACC, sizeof(*c->qs_tran));
  c->qs_freq = calloc(NACC, sizeof(*c->qs_freq));
  c->er_init = calloc(NACC, sizeof(*c->er_init));
  c->er_emis = calloc(NACC, sizeof(*c->er_emis));
  c->er_tran = calloc(NACC, sizeof(*c->er_tran));
  c->freq_acc = (int64_t *)calloc(100001, sizeof(int64_t));
  strcpy(c->id_prefix, "S");
  /* ref: :546-549  qc[i].prob = pow(10, (double)i / -10) */
  for (i = 0; i <= 93; i++) c->qc_prob[i] = pow(10, (double)i / -10);
  /* ref: :558-578  uniform error probability: two adjacent QVs whose mean error is 1-acc/100 */
  for (i = 0; i <= ORC_ACC_MAX; i++) {
    for (j = 0; j <= 93; j++) c->uni_ep[i][j] = 0;
    if (i == ORC_ACC_MAX) {
      c->uni_ep[i][93] = 1.0;
      continue;
    }
    prob = 1.0 - i / 100.0;
    for (j = 0; j <= 93; j++) {
      if (prob == c->qc_prob[j]) {
        c->uni_ep[i][j] = 1.0;
        break;
      } else if (prob > c->qc_prob[j]) {
        rate = (prob - c->qc_prob[j]) / (c->qc_prob[j - 1] - c->qc_prob[j]);
        c->uni_ep[i][j - 1] = rate;
        c->uni_ep[i][j] = 1 - rate;
        break;
      }
    }
  }
  for (i = 1; i <= 10; i++) c->bias[i] = 1; /* ref: :673-676 */
  return c;
}

void orc_free(orc_ctx *c) {
  if (!c) return;
  free(c->ip); free(c->ep); free(c->tp);
  free(c->prob2len); free(c->prob2acc);
  free(c->qs_init); free(c->qs_emis); free(c->qs_tran); free(c->qs_freq);
  free(c->er_init); free(c->er_emis); free(c->er_tran);
  free(c->seq); free(c->hp_alloc);
  free(c->w_seq); free(c->read_seq); free(c->qual); free(c->maf_seq); free(c->maf_ref); free(c->w_hp_alloc);
  free(c->rng.rec);
  free(c->out_reads.p); free(c->out_maf.p);
  free(c->freq_len); free(c->freq_acc); free(c->info);
  free(c);
}

const char *orc_error(orc_ctx *c) { return c->err; }

void orc_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  orc_philox4x32_10(ctr, key, out);
}

/* ref: the C++ functional cast int(x) on a double; x86-64 cvttsd2si yields INT_MIN for
 * NaN / out-of-range, which is what the reference build observes. */
static long trunc_int(double x) {
  if (!(x > -2147483649.0 && x < 2147483648.0)) return (long)INT_MIN;
  return (long)(int)x;
}

int orc_set_params(orc_ctx *c, int method, int pass_num, double accuracy_mean,
                   long len_min, long len_max, double len_mean, double len_sd,
                   long sub_ratio, long ins_ratio, long del_ratio,
                   double hp_del_bias, const char *id_prefix) {
  long sum;
  int i;
  c->method = method;
  c->pass_num = pass_num;
  c->accuracy_mean = accuracy_mean;
  c->len_min = len_min;
  c->len_max = len_max;
  c->len_mean = len_mean;
  c->len_sd = len_sd;
  c->sub_ratio = sub_ratio;
  c->ins_ratio = ins_ratio;
  c->del_ratio = del_ratio;
  c->hp_del_bias_opt = hp_del_bias;
  snprintf(c->id_prefix, sizeof c->id_prefix, "%s", id_prefix ? id_prefix : "S");
  /* ref: :1561-1564 */
  sum = sub_ratio + ins_ratio + del_ratio;
  c->sub_rate = (double)sub_ratio / sum;
  c->ins_rate = (double)ins_ratio / sum;
  c->del_rate = (double)del_ratio / sum;
  /* ref: set_mut :5474-5479 */
  for (i = 0; i <= 93; i++) {
    c->sub_thre[i] = trunc_int((c->qc_prob[i] * c->sub_rate) * 1000000 + 0.5);
    c->ins_thre[i] = trunc_int((c->qc_prob[i] * (c->sub_rate + c->ins_rate)) * 1000000 + 0.5);
    c->del_thre[i] = trunc_int((c->qc_prob[i] * c->del_rate) / (1 + c->qc_prob[i] * c->del_rate) * 1000000 + 0.5);
  }
  free(c->freq_len);
  c->freq_len_n = 2 * len_max + 2; /* the reference indexes freq_len[len] unguarded (:2300) */
  c->freq_len = (int64_t *)calloc((size_t)c->freq_len_n, sizeof(int64_t));
  c->tables_built = 0;
  return 0;
}

/* ref: set_qshmm :5570-5634, set_errhmm :5640-5714.  Format:
 *   <acc> IP <state> <p> | <acc> EP <state> <p...> (0-based columns) | <acc> TP <state> <p...> (1-based) */
int orc_load_model(orc_ctx *c, const char *path) {
  FILE *fp = fopen(path, "r");
  char *line, *tok;
  int accuracy, state, num, i;
  int64_t ep_cells, tp_cells, ip_cells, idx;
  if (!fp) return fail(c, "ERROR: Cannot open file (model)");
  c->ep_cols = (c->method == ORC_METHOD_ERR) ? 4 : 94;
  free(c->ep);
  ep_cells = (int64_t)NACC * NST * c->ep_cols;
  tp_cells = (int64_t)NACC * NST * NST;
  ip_cells = (int64_t)NACC * NST;
  c->ep = (double *)calloc((size_t)ep_cells, sizeof(double));
  memset(c->ip, 0, (size_t)ip_cells * sizeof(double));
  memset(c->tp, 0, (size_t)tp_cells * sizeof(double));
  for (i = 0; i < NACC; i++) {
    c->exist[i] = 0;
    c->state_max[i] = 0;
  }
  c->acc_min = 100;
  c->acc_max = 0;
  line = (char *)malloc(LINE_MAX_BYTES);
  while (fgets(line, LINE_MAX_BYTES, fp) != NULL) {
    size_t n = strlen(line);
    if (n && line[n - 1] == '\n') line[n - 1] = '\0';
    tok = strtok(line, " ");
    if (!tok) continue;
    accuracy = atoi(tok);
    if (accuracy < 0 || accuracy > ORC_ACC_MAX) { fclose(fp); free(line); return fail(c, "model: accuracy out of range"); }
    c->exist[accuracy] = 1;
    if (c->acc_min > accuracy) c->acc_min = accuracy;
    if (c->acc_max < accuracy) c->acc_max = accuracy;
    tok = strtok(NULL, " ");
    if (!tok) continue;
    if (strcmp(tok, "IP") == 0) {
      tok = strtok(NULL, " ");
      state = atoi(tok);
      tok = strtok(NULL, " ");
      idx = (int64_t)accuracy * NST + state;
      if (idx < 0 || idx >= ip_cells) { fclose(fp); free(line); return fail(c, "model: IP index outside the reference array"); }
      c->ip[idx] = atof(tok);
      c->state_max[accuracy] = state; /* ref: :5686 (errhmm only uses it) */
    } else if (strcmp(tok, "EP") == 0) {
      tok = strtok(NULL, " ");
      state = atoi(tok);
      num = 0;
      tok = strtok(NULL, " ");
      while (tok != NULL) {
        idx = ((int64_t)accuracy * NST + state) * c->ep_cols + num;
        if (idx < 0 || idx >= ep_cells) { fclose(fp); free(line); return fail(c, "model: EP index outside the reference array"); }
        c->ep[idx] = atof(tok);
        num++;
        tok = strtok(NULL, " ");
      }
    } else if (strcmp(tok, "TP") == 0) {
      tok = strtok(NULL, " ");
      state = atoi(tok);
      num = 0;
      tok = strtok(NULL, " ");
      while (tok != NULL) {
        num++;
        idx = ((int64_t)accuracy * NST + state) * NST + num;
        if (idx < 0 || idx >= tp_cells) { fclose(fp); free(line); return fail(c, "model: TP index outside the reference array"); }
        c->tp[idx] = atof(tok);
        tok = strtok(NULL, " ");
      }
    }
  }
  free(line);
  fclose(fp);
  c->model_loaded = 1;
  c->tables_built = 0;
  return 0;
}

#define IP(a, s) c->ip[(int64_t)(a) * NST + (s)]
#define EP(a, s, k) c->ep[((int64_t)(a) * NST + (s)) * c->ep_cols + (k)]
#define TP(a, s, k) c->tp[((int64_t)(a) * NST + (s)) * NST + (k)]

/* ref: the table builders inlined at the top of simulate_by_qshmm (:1991-2170) and
 * simulate_by_errhmm (:3633-3789).  Quantised CDFs: walk outcomes in index order, skip
 * zero-probability outcomes, end = int(cum*R + 0.5) clamped to R, fill (start..end],
 * stop once end >= R; the row modulus is the last `end` (stale when a row is empty, as
 * in the reference, because end_wk is one function-scope variable). */
int orc_build_tables(orc_ctx *c) {
  long i, j, k, l;
  double variance, kappa, theta, gam, mean;
  double len_prob_total, freq_total, accuracy_prob_total, cum;
  long start_wk, end_wk = 0;
  long accuracy_min, accuracy_max;

  if (!c->model_loaded) return fail(c, "model not loaded");

  /* length distribution (ref: :1991-2027) */
  variance = pow(c->len_sd, 2);
  kappa = pow(c->len_mean, 2) / variance;
  theta = variance / c->len_mean;
  gam = tgamma(kappa);
  if (c->len_sd == 0.0) {
    c->prob2len[1] = trunc_int(c->len_mean + 0.5);
    c->len_rand_value = 1;
  } else {
    start_wk = 1;
    len_prob_total = 0.0;
    for (i = c->len_min; i <= c->len_max; i++) {
      len_prob_total += pow((double)i, kappa - 1) * exp((double)(-1 * i) / theta) / pow(theta, kappa) / gam;
      end_wk = trunc_int(len_prob_total * 100000 + 0.5);
      if (end_wk > 100000) end_wk = 100000;
      for (j = start_wk; j <= end_wk; j++) c->prob2len[j] = i;
      if (end_wk >= 100000) break;
      start_wk = end_wk + 1;
    }
    c->len_rand_value = end_wk;
  }
  /* the reference makes this check only with --pass-num 1 (:2022, :3664); with more passes it goes on and indexes
   * prob2len with a draw modulo a non-positive value (a crash in practice).  The restatement stops in both cases, as
   * the product's table builder does. */
  if (c->len_rand_value < 1)
    return fail(c, "ERROR: length parameters are not appropriate.");

  /* accuracy distribution (ref: :2029-2064) */
  mean = c->accuracy_mean * 100;
  accuracy_max = (long)floor(mean * 1.05);
  accuracy_min = (long)floor(mean * 0.75);
  if (accuracy_max > 100) accuracy_max = 100;
  freq_total = 0.0;
  for (i = accuracy_min; i <= accuracy_max; i++) freq_total += exp(0.22 * i);
  start_wk = 1;
  accuracy_prob_total = 0.0;
  for (i = accuracy_min; i <= accuracy_max; i++) {
    accuracy_prob_total += exp(0.22 * i) / freq_total;
    end_wk = trunc_int(accuracy_prob_total * 100000 + 0.5);
    if (end_wk > 100000) end_wk = 100000;
    for (j = start_wk; j <= end_wk; j++) c->prob2acc[j] = i;
    if (end_wk >= 100000) break;
    start_wk = end_wk + 1;
  }
  c->accuracy_rand_value = end_wk;
  if (c->accuracy_rand_value < 1) return fail(c, "ERROR: accuracy parameters are not appropriate.");
  c->tab_acc_lo = accuracy_min;
  c->tab_acc_hi = accuracy_max;

  if (c->method == ORC_METHOD_QS) {
    /* ref: :2066-2170 */
    for (i = accuracy_min; i <= accuracy_max; i++) {
      if (c->exist[i] == 1) {
        start_wk = 1;
        cum = 0.0;
        for (j = 1; j <= ORC_STATE_MAX; j++) {
          if (IP(i, j) == 0) continue;
          cum += IP(i, j);
          end_wk = trunc_int(cum * 100 + 0.5);
          if (end_wk > 100) end_wk = 100;
          for (k = start_wk; k <= end_wk; k++) c->qs_init[i][k] = (uint8_t)j;
          if (end_wk >= 100) break;
          start_wk = end_wk + 1;
        }
        c->mod_init[i] = end_wk;
        for (j = 1; j <= ORC_STATE_MAX; j++) {
          start_wk = 1;
          cum = 0.0;
          for (k = 0; k <= 93; k++) {
            if (EP(i, j, k) == 0) continue;
            cum += EP(i, j, k);
            end_wk = trunc_int(cum * 100 + 0.5);
            if (end_wk > 100) end_wk = 100;
            for (l = start_wk; l <= end_wk; l++) c->qs_emis[i][j][l] = (uint8_t)k;
            if (end_wk >= 100) break;
            start_wk = end_wk + 1;
          }
          c->mod_emis[i][j] = end_wk;
        }
        for (j = 1; j <= ORC_STATE_MAX; j++) {
          start_wk = 1;
          cum = 0.0;
          for (k = 1; k <= ORC_STATE_MAX; k++) {
            if (TP(i, j, k) == 0) continue;
            cum += TP(i, j, k);
            end_wk = trunc_int(cum * 100 + 0.5);
            if (end_wk > 100) end_wk = 100;
            for (l = start_wk; l <= end_wk; l++) c->qs_tran[i][j][l] = (uint8_t)k;
            if (end_wk >= 100) break;
            start_wk = end_wk + 1;
          }
          c->mod_tran[i][j] = end_wk;
        }
      } else {
        start_wk = 1;
        cum = 0.0;
        for (j = 0; j <= 93; j++) {
          if (c->uni_ep[i][j] == 0) continue;
          cum += c->uni_ep[i][j];
          end_wk = trunc_int(cum * 1000 + 0.5);
          if (end_wk > 1000) end_wk = 1000;
          for (k = start_wk; k <= end_wk; k++) c->qs_freq[i][k] = (uint8_t)j;
          if (end_wk >= 1000) break;
          start_wk = end_wk + 1;
        }
        c->mod_freq[i] = end_wk;
      }
    }
  } else {
    /* ref: :3708-3789 */
    for (i = accuracy_min; i <= accuracy_max; i++) {
      if (c->exist[i] == 0) continue;
      start_wk = 1;
      cum = 0.0;
      for (j = 1; j <= c->state_max[i]; j++) {
        if (IP(i, j) == 0) continue;
        cum += IP(i, j);
        end_wk = trunc_int(cum * 1000 + 0.5);
        if (end_wk > 1000) end_wk = 1000;
        for (k = start_wk; k <= end_wk; k++) c->er_init[i][k] = (uint8_t)j;
        if (end_wk >= 1000) break;
        start_wk = end_wk + 1;
      }
      c->mod_init[i] = end_wk;
      for (j = 1; j <= c->state_max[i]; j++) {
        start_wk = 1;
        cum = 0.0;
        c->er_del[i][j] = trunc_int(EP(i, j, 3) * 1000 + 0.5);
        for (k = 0; k <= 2; k++) {
          if (EP(i, j, k) <= 0) continue;
          cum += EP(i, j, k);
          end_wk = trunc_int(cum * 1000 + 0.5);
          if (end_wk > 1000) end_wk = 1000;
          for (l = start_wk; l <= end_wk; l++) c->er_emis[i][j][l] = (uint8_t)k;
          if (end_wk >= 1000) break;
          start_wk = end_wk + 1;
        }
        c->mod_emis[i][j] = end_wk;
      }
      for (j = 1; j <= c->state_max[i]; j++) {
        start_wk = 1;
        cum = 0.0;
        for (k = 1; k <= ORC_STATE_MAX; k++) {
          if (TP(i, j, k) == 0) continue;
          cum += TP(i, j, k);
          end_wk = trunc_int(cum * 1000 + 0.5);
          if (end_wk > 1000) end_wk = 1000;
          for (l = start_wk; l <= end_wk; l++) c->er_tran[i][j][l] = (uint8_t)k;
          if (end_wk >= 1000) break;
          start_wk = end_wk + 1;
        }
        c->mod_tran[i][j] = end_wk;
      }
    }
  }
  c->tables_built = 1;
  return 0;
}

/* ------------------------------------------------------------------ genome */

/* ref: get_genome_seq :1035-1065 — upper-case, then per-base homopolymer length with the
 * reference's counter quirk (nnum>11 -> 10, so runs >= 11 alternate 11,10,11,...), N runs
 * get hp=1.  hpfreq[nnum]++ with nnum==11 writes one past hpfreq[11], which in the
 * reference build (g++ 13.3 -O2, struct genome_t :93-103) is hp_del_bias[0]; we keep the
 * same aliasing in hpfreq[11] and derive bias[0] from it in refresh_bias0(). */
/* upper_from: first index toupper() is applied to.  get_genome_seq (:1035) and simulate_by_qshmm_trans (:2774)
 * start at 0; simulate_by_qshmm_templ (:3330), simulate_by_errhmm_trans (:4474) and simulate_by_errhmm_templ
 * loop i = 1..len, so the FIRST base of the sequence keeps its case there.
 * weight: transcripts count every run read_num times in the bias prepass (:2714-2718). */
static void hp_scan_w(orc_ctx *c, char *seq, int64_t len, int16_t *hp, int upper_from, long weight, int count) {
  int64_t i, j, nstart = 0, nend = 0;
  int16_t nnum = 1;
  for (i = upper_from; i < len; i++) {
    char ch = seq[i];
    if (ch >= 'a' && ch <= 'z') seq[i] = (char)(ch - 'a' + 'A'); /* toupper in the C locale */
  }
  for (i = 1; i <= len; i++) {
    if ((i < len) && (seq[i - 1] == seq[i])) {
      nend = i;
      nnum++;
      if (nnum > 11) nnum = 10;
    } else {
      int bin = (seq[i - 1] == 'N') ? 1 : nnum;
      for (j = nstart; j <= nend; j++) {
        if (hp) hp[j] = (int16_t)bin;
      }
      if (count) c->hpfreq[bin] += weight * (long)(nend - nstart + 1);
      nstart = i;
      nend = nstart;
      nnum = 1;
    }
  }
}

static void hp_scan(orc_ctx *c, char *seq, int64_t len, int16_t *hp) { hp_scan_w(c, seq, len, hp, 0, 1, 1); }

/* bias[0] is the double whose bit pattern is the long hpfreq[11] (see above); bias[11] reads
 * the zero padding after `genome` in the reference build -> 0.0. */
static void refresh_bias0(orc_ctx *c) {
  int64_t bits = (int64_t)c->hpfreq[11];
  memcpy(&c->bias[0], &bits, sizeof(double));
  c->bias[11] = 0.0;
}

int orc_prepass_sequence(orc_ctx *c, const char *seq, int64_t len) {
  char *tmp = (char *)malloc((size_t)len + 1);
  memcpy(tmp, seq, (size_t)len);
  tmp[len] = 0;
  hp_scan(c, tmp, len, NULL);
  free(tmp);
  return 0;
}

/* ref: main :678-697.  Called before the first prepass_sequence it zeroes hpfreq[0..10]
 * (not [11]); called with finish=1 it computes the normalised bias. */
int orc_finish_bias(orc_ctx *c) {
  long sum1 = 0, sum2 = 0;
  double rate;
  int i;
  for (i = 1; i <= 10; i++) {
    c->bias[i] = 1 + (c->hp_del_bias_opt - 1) / 9 * (i - 1);
    sum1 += c->hpfreq[i] * c->bias[i];
    sum2 += c->hpfreq[i];
  }
  rate = (double)sum2 / sum1;
  for (i = 1; i <= 10; i++) c->bias[i] *= rate;
  refresh_bias0(c);
  return 0;
}

int orc_set_sequence(orc_ctx *c, const char *seq, int64_t len, int seq_num) {
  free(c->seq);
  free(c->hp_alloc);
  c->seq = (char *)malloc((size_t)len + 1);
  memcpy(c->seq, seq, (size_t)len);
  c->seq[len] = 0;
  c->hp_alloc = (int16_t *)calloc((size_t)len + 2, sizeof(int16_t));
  c->hp = c->hp_alloc + 1;
  c->glen = len;
  c->seq_num = seq_num;
  hp_scan(c, c->seq, len, c->hp);
  refresh_bias0(c);
  return 0;
}

void orc_get_bias(orc_ctx *c, double bias[12]) { memcpy(bias, c->bias, sizeof c->bias); }
const int16_t *orc_get_hp(orc_ctx *c, int64_t *n) { *n = c->glen; return c->hp; }
const char *orc_get_seq(orc_ctx *c, int64_t *n) { *n = c->glen; return c->seq; }

/* ------------------------------------------------------------------ draw source setup */
int orc_rng_glibc(orc_ctx *c, uint32_t seed) {
  c->rng.mode = RNG_GLIBC;
  orc_glibc_srand(&c->rng.g, seed);
  c->rng.cur = 0;
  c->rng.exhausted = 0;
  return 0;
}
int orc_rng_replay(orc_ctx *c, const int32_t *log, int64_t n) {
  c->rng.mode = RNG_REPLAY;
  c->rng.log = log;
  c->rng.nlog = n;
  c->rng.cur = 0;
  c->rng.exhausted = 0;
  return 0;
}
int orc_rng_philox(orc_ctx *c, uint32_t seed) {
  c->rng.mode = RNG_PHILOX;
  c->rng.key[0] = seed;
  c->rng.key[1] = 0;
  c->rng.cur = 0;
  c->rng.exhausted = 0;
  return 0;
}

/* ------------------------------------------------------------------ helpers */

/* ref: count_digit :5823-5835 */
static int count_digit(long num) {
  int digit = 1;
  int quotient = (int)(num / 10);
  while (quotient != 0) {
    digit++;
    quotient = (int)(quotient / 10);
  }
  return digit;
}

/* ref: revcomp :5841-5864 — reverse, then complement upper-case A/T/G/C only */
static void revcomp_n(char *s, int64_t len) {
  int64_t i;
  for (i = 0; i < len / 2; i++) {
    char t = s[i];
    s[i] = s[len - i - 1];
    s[len - i - 1] = t;
  }
  for (i = 0; i < len; i++) {
    switch (s[i]) {
      case 'A': s[i] = 'T'; break;
      case 'T': s[i] = 'A'; break;
      case 'G': s[i] = 'C'; break;
      case 'C': s[i] = 'G'; break;
      default: break;
    }
  }
}

static void ensure_window(orc_ctx *c, int64_t wlen) {
  int64_t cap = 4 * wlen + 4096; /* the reference uses 2*len_max+1 and overflows silently */
  if (cap <= c->w_cap) return;
  c->w_cap = cap;
  c->w_seq = (char *)realloc(c->w_seq, (size_t)cap);
  c->read_seq = (char *)realloc(c->read_seq, (size_t)cap);
  c->qual = (char *)realloc(c->qual, (size_t)cap);
  c->maf_seq = (char *)realloc(c->maf_seq, (size_t)cap);
  c->maf_ref = (char *)realloc(c->maf_ref, (size_t)cap);
  free(c->w_hp_alloc);
  c->w_hp_alloc = (int16_t *)calloc((size_t)cap + 1, sizeof(int16_t));
  c->w_hp = c->w_hp_alloc + 1; /* w_hp[-1] == 0: ref reads mut.hp[-1] (malloc header, 0) :2269 */
}

static const char SUB_A[] = "TGC", SUB_T[] = "AGC", SUB_G[] = "ATC", SUB_C[] = "ATG", NT4[] = "ATGC"; /* ref: :5481-5486 */

static char substitute(rng_t *r, char nt) {
  uint32_t index = d_choice3(r); /* drawn even when the base is not ACGT (ref: :2235-2246) */
  switch (nt) {
    case 'A': return SUB_A[index];
    case 'T': return SUB_T[index];
    case 'G': return SUB_G[index];
    case 'C': return SUB_C[index];
    default: return NT4[d_choice4(r)];
  }
}

static void push_info(orc_ctx *c, const orc_readinfo_t *ri) {
  if (c->info_n >= c->info_cap) {
    c->info_cap = c->info_cap ? c->info_cap * 2 : 1024;
    c->info = (orc_readinfo_t *)realloc(c->info, (size_t)c->info_cap * sizeof(orc_readinfo_t));
  }
  c->info[c->info_n++] = *ri;
}

/* ref: record emission :2318-2383 (= :4012-4078): FASTQ or SAM, then MAF */
static void emit_records(orc_ctx *c, long read_num, long pass, long offset, long wlen, char strand,
                         long len, long ncol) {
  char id[512];
  int d1[4], d2[4], dn[4], i;
  buf_t *o = &c->out_reads, *m = &c->out_maf;
  const char *name = c->set_mode ? c->rec_name : "ref";
  if (c->pass_num == 1) {
    if (c->set_mode) snprintf(id, sizeof id, "%s_%ld", c->id_prefix, read_num);
    else snprintf(id, sizeof id, "%s%d_%ld", c->id_prefix, c->seq_num, read_num);
    buf_puts(o, "@"); buf_puts(o, id); buf_puts(o, "\n");
    buf_put(o, c->read_seq, len);
    buf_puts(o, "\n+"); buf_puts(o, id); buf_puts(o, "\n");
    buf_put(o, c->qual, len);
    buf_puts(o, "\n");
  } else {
    char tail[256];
    if (c->set_mode) snprintf(id, sizeof id, "%s/%ld/%ld", c->id_prefix, read_num, pass);
    else snprintf(id, sizeof id, "%s%d/%ld/%ld", c->id_prefix, c->seq_num, read_num, pass);
    buf_puts(o, id);
    buf_puts(o, "\t4\t*\t0\t255\t*\t*\t0\t0\t");
    buf_put(o, c->read_seq, len);
    buf_puts(o, "\t");
    buf_put(o, c->qual, len);
    buf_puts(o, "\tcx:i:3\tip:B:C");
    for (i = 0; i < len; i++) buf_put(o, ",9", 2);
    buf_puts(o, "\tnp:i:1\tpw:B:C");
    for (i = 0; i < len; i++) buf_put(o, ",9", 2);
    snprintf(tail, sizeof tail, "\tqs:i:0\tqe:i:%ld\trq:f:%f\tsn:B:f,10.0,10.0,10.0,10.0\tzm:i:%ld\tRG:Z:ffffffff\n",
             (long)(int)(len - 1), c->accuracy_mean, read_num);
    buf_puts(o, tail);
  }
  d1[0] = c->set_mode ? c->rec_name_width : 3;
  d2[0] = 1 + count_digit(read_num);
  d1[1] = count_digit(offset);    d2[1] = 1;
  d1[2] = count_digit(wlen);      d2[2] = count_digit(len);
  d1[3] = count_digit(c->glen);   d2[3] = count_digit(len);
  for (i = 0; i < 4; i++) dn[i] = d1[i] >= d2[i] ? d1[i] : d2[i];
  buf_puts(m, "a\ns ");
  buf_puts(m, name);
  buf_pad(m, dn[0] - d1[0]);
  buf_pad(m, dn[1] - d1[1]);
  buf_long(m, " ", offset, "");
  buf_pad(m, dn[2] - d1[2]);
  buf_long(m, " ", wlen, " +");
  buf_pad(m, dn[3] - d1[3]);
  buf_long(m, " ", (long)c->glen, " ");
  buf_put(m, c->maf_ref, ncol);
  buf_puts(m, "\ns ");
  buf_puts(m, id);
  buf_pad(m, dn[0] - d2[0]);
  buf_pad(m, dn[1] - d2[1]);
  buf_puts(m, " 0");
  buf_pad(m, dn[2] - d2[2]);
  buf_long(m, " ", len, strand == '+' ? " +" : " -");
  buf_pad(m, dn[3] - d2[3]);
  buf_long(m, " ", len, " ");
  buf_put(m, c->maf_seq, ncol);
  buf_puts(m, "\n\n");
}

/* ------------------------------------------------------------------ per-pass generators */

typedef struct {
  long rlen, ncol, nsub, nins, ndel;
} pass_out_t;

/* ref: simulate_by_qshmm inner loops :2210-2286 */
static void qshmm_pass(orc_ctx *c, rng_t *r, uint32_t pass, int acc, long wlen, pass_out_t *po) {
  long ref_offset = 0, read_offset = 0, maf_offset = 0;
  long state = 0, index, qv, rand_value;
  char nt;
  po->nsub = po->nins = po->ndel = 0;
  while (ref_offset < wlen) {
    d_begin(r, pass, (uint32_t)read_offset);
    if (c->exist[acc] == 1) {
      if (read_offset == 0) {
        index = d_state(r, (uint32_t)c->mod_init[acc]) + 1;
        state = c->qs_init[acc][index];
      } else {
        index = d_state(r, (uint32_t)c->mod_tran[acc][state]) + 1;
        state = c->qs_tran[acc][state][index];
      }
      index = d_qs_emis(r, (uint32_t)c->mod_emis[acc][state]) + 1;
      qv = c->qs_emis[acc][state][index];
    } else {
      index = d_qs_freq(r, (uint32_t)c->mod_freq[acc]) + 1;
      qv = c->qs_freq[acc][index];
    }
    c->qual[read_offset] = (char)(qv + 33);
    nt = c->w_seq[ref_offset];
    rand_value = d_qs_err(r);
    if (d_lt(r, (uint32_t)rand_value, (double)c->sub_thre[qv])) {
      po->nsub++;
      c->read_seq[read_offset] = substitute(r, nt);
      c->maf_ref[maf_offset] = nt;
      ref_offset++;
    } else if (d_lt(r, (uint32_t)rand_value, (double)c->ins_thre[qv])) {
      po->nins++;
      index = d_choice8(r);
      c->read_seq[read_offset] = (index >= 4) ? nt : NT4[index];
      c->maf_ref[maf_offset] = '-';
    } else {
      c->read_seq[read_offset] = nt;
      c->maf_ref[maf_offset] = nt;
      ref_offset++;
    }
    c->maf_seq[maf_offset] = c->read_seq[read_offset];
    maf_offset++;
    read_offset++;
    {
      uint32_t j = 0;
      while (ref_offset < wlen) {
        int hp = c->w_hp[ref_offset - 1];
        rand_value = d_del(r, j++);
        if (d_lt(r, (uint32_t)rand_value, c->del_thre[qv] * c->bias[hp])) {
          po->ndel++;
          c->maf_seq[maf_offset] = '-';
          c->maf_ref[maf_offset] = c->w_seq[ref_offset];
          maf_offset++;
          ref_offset++;
        } else {
          break;
        }
      }
    }
  }
  po->rlen = read_offset;
  po->ncol = maf_offset;
}

/* ref: simulate_by_errhmm inner loops :3836-3976 */
static void errhmm_pass(orc_ctx *c, rng_t *r, uint32_t pass, int acc, int rate_mag, long wlen, pass_out_t *po) {
  long ref_offset = 0, read_offset = 0, maf_offset = 0, i;
  long state = 0, index, index2;
  int tacc, hp;
  char nt;
  po->nsub = po->nins = po->ndel = 0;
  if (acc == 100) {
    for (i = 0; i < wlen; i++) {
      nt = c->w_seq[i];
      c->read_seq[i] = nt;
      c->maf_ref[i] = nt;
      c->maf_seq[i] = nt;
    }
    po->rlen = wlen;
    po->ncol = wlen;
    return;
  }
  /* which accuracy's tables drive the chain (ref: :3852, :3872, :3900) */
  if (c->exist[acc] == 1) tacc = acc;
  else if (acc < c->acc_min) tacc = c->acc_min;
  else tacc = c->acc_max;
  while (ref_offset < wlen) {
    nt = c->w_seq[ref_offset];
    d_begin(r, pass, (uint32_t)maf_offset);
    if (read_offset == 0) {
      index = d_state(r, (uint32_t)c->mod_init[tacc]) + 1;
      state = c->er_init[tacc][index];
    } else {
      index = d_state(r, (uint32_t)c->mod_tran[tacc][state]) + 1;
      state = c->er_tran[tacc][state][index];
    }
    hp = c->w_hp[ref_offset];
    index = d_w1(r, 1000) + 1;
    if (index <= c->er_del[tacc][state] * c->bias[hp]) {
      index = 3;
    } else {
      if (c->mod_emis[tacc][state] == 0) {
        index = d_w2(r, 3);
      } else {
        index = d_w2(r, (uint32_t)c->mod_emis[tacc][state]) + 1;
        index = c->er_emis[tacc][state][index];
      }
    }
    if (c->exist[acc] != 1) {
      if (acc < c->acc_min) {
        if (index == 0) { /* ref: :3892-3899 thicken errors */
          index = d_w3(r, 100) + 1;
          if (index <= rate_mag) index = d_mag3(r);
          else index = 0;
        }
      } else {
        if (index != 0) { /* ref: :3920-3925 thin errors */
          index2 = d_w3(r, 100) + 1;
          if (index2 <= rate_mag) index = 0;
        }
      }
    }
    if (index == 0) {
      c->read_seq[read_offset] = nt;
      c->maf_seq[maf_offset] = nt;
      c->maf_ref[maf_offset] = nt;
      ref_offset++;
      read_offset++;
    } else if (index == 1) {
      po->nsub++;
      c->read_seq[read_offset] = substitute(r, nt);
      c->maf_seq[maf_offset] = c->read_seq[read_offset];
      c->maf_ref[maf_offset] = nt;
      ref_offset++;
      read_offset++;
    } else if (index == 2) {
      po->nins++;
      index = d_choice8(r);
      c->read_seq[read_offset] = (index >= 4) ? nt : NT4[index];
      c->maf_seq[maf_offset] = c->read_seq[read_offset];
      c->maf_ref[maf_offset] = '-';
      read_offset++;
    } else {
      po->ndel++;
      c->maf_seq[maf_offset] = '-';
      c->maf_ref[maf_offset] = nt;
      ref_offset++;
    }
    maf_offset++;
  }
  po->rlen = read_offset;
  po->ncol = maf_offset;
}

/* ------------------------------------------------------------------ WGS driver */

void orc_reset_outputs(orc_ctx *c) {
  c->out_reads.n = 0;
  c->out_maf.n = 0;
  c->info_n = 0;
}

/* copies the window [offset, offset+wlen) of the current sequence, reverse-complemented for '-'
 * (ref: :2193-2207 = :2874-2886) */
static void load_window(orc_ctx *c, long offset, long wlen, char strand) {
  long i;
  ensure_window(c, wlen);
  for (i = 0; i < wlen; i++) {
    c->w_seq[i] = c->seq[offset + i];
    c->w_hp[i] = c->hp[offset + i];
  }
  c->w_seq[wlen] = '\0';
  if (strand == '-') {
    revcomp_n(c->w_seq, wlen);
    for (i = 0; i < wlen / 2; i++) { /* ref: revshort :5870-5879 */
      int16_t t = c->w_hp[i];
      c->w_hp[i] = c->w_hp[wlen - i - 1];
      c->w_hp[wlen - i - 1] = t;
    }
  }
}

/* PHILOX mode: sum of the error probabilities of qual[0..len) in 2^-26 fixed point (order independent) */
static double qs_prob_sum_fixed(orc_ctx *c, long len) {
  uint64_t acc = 0;
  long i;
  for (i = 0; i < len; i++) acc += (uint64_t)llround(c->qc_prob[(int)c->qual[i] - 33] * 67108864.0);
  return (double)acc / 67108864.0;
}

/* all passes of one read: chains, per-read statistics, records (ref: :2209-2383 = :2888-3017 = :3366-3530,
 * errhmm :3836-4078).  *len_total_pass0 receives the emitted length of pass 0 (the WGS quota counter). */
static void run_read_passes(orc_ctx *c, long offset, long wlen, char strand, int acc, int64_t start_draw,
                            long *len_pass0, double *accuracy_total) {
  rng_t *r = &c->rng;
  orc_stats_t *st = &c->st;
  int rate_mag = 0;
  long h, i, len;
  double value;
  if (c->method == ORC_METHOD_ERR) { /* ref: :3829-3833 */
    if (acc < c->acc_min) rate_mag = (int)((double)(c->acc_min - acc) / c->acc_min * 100);
    else if (acc > c->acc_max) rate_mag = (int)((double)(acc - c->acc_max) / (100 - c->acc_max) * 100);
  }
  for (h = 0; h < c->pass_num; h++) {
    pass_out_t po;
    orc_readinfo_t ri;
    int64_t pass_start = (h == 0) ? start_draw : r->cur;
    if (c->method == ORC_METHOD_QS) qshmm_pass(c, r, (uint32_t)h, acc, wlen, &po);
    else errhmm_pass(c, r, (uint32_t)h, acc, rate_mag, wlen, &po);
    len = po.rlen;
    if (strand == '-') {
      revcomp_n(c->maf_seq, po.ncol);
      revcomp_n(c->maf_ref, po.ncol);
    }
    st->res_sub_num += po.nsub;
    st->res_ins_num += po.nins;
    st->res_del_num += po.ndel;
    st->res_len_total += len;
    if (h == 0) *len_pass0 = len;
    if (len >= 0 && len < c->freq_len_n) c->freq_len[len]++;
    if (len > st->res_len_max) st->res_len_max = len;
    if (len < st->res_len_min) st->res_len_min = len;
    if (c->method == ORC_METHOD_QS) { /* ref: :2309-2316 accuracy from emitted qualities */
      double prob = 0.0;
      if (r->mode == RNG_PHILOX) {
        /* engine definition for PHILOX mode: the error probabilities are summed in fixed point (26 fractional
         * bits), so the sum does not depend on the order in which the position-parallel pass 1 adds them */
        prob = qs_prob_sum_fixed(c, len);
      } else {
        for (i = 0; i < len; i++) prob += c->qc_prob[(int)c->qual[i] - 33];
      }
      value = 1.0 - (prob / len);
    } else { /* ref: :4002 accuracy from realised errors; qualities all '!' :4007-4010 */
      value = 1.0 - ((double)(po.nsub + po.nins + po.ndel) / len);
      for (i = 0; i < len; i++) c->qual[i] = '!';
    }
    *accuracy_total += value;
    {
      long acc_wk = trunc_int(value * 100000 + 0.5);
      if (acc_wk >= 0 && acc_wk <= 100000) c->freq_acc[acc_wk]++;
    }
    emit_records(c, (long)st->res_num, h, offset, wlen, strand, len, po.ncol);
    ri.read_id = st->res_num; ri.pass = (int32_t)h; ri.acc = acc; ri.offset = offset; ri.wlen = wlen;
    ri.rlen = len; ri.ncol = po.ncol; ri.strand = strand; ri.nsub = (int32_t)po.nsub;
    ri.nins = (int32_t)po.nins; ri.ndel = (int32_t)po.ndel; ri.draw_start = pass_start; ri.accuracy = value;
    push_info(c, &ri);
  }
}

static void begin_stats(orc_ctx *c) {
  orc_stats_t *st = &c->st;
  memset(st, 0, sizeof *st);
  st->res_len_min = LONG_MAX;
  memset(c->freq_len, 0, (size_t)c->freq_len_n * sizeof(int64_t));
  memset(c->freq_acc, 0, 100001 * sizeof(int64_t));
}

/* ref: :2387-2410 (= :3023-3047) and print_simulation_stats :5543, :5557-5559 */
static void finish_stats(orc_ctx *c, double accuracy_total, int64_t depth_len) {
  orc_stats_t *st = &c->st;
  double variance;
  long i;
  st->res_pass_num = st->res_num * c->pass_num;
  st->res_len_mean = (double)st->res_len_total / st->res_pass_num;
  st->res_accuracy_mean = accuracy_total / st->res_pass_num;
  st->accuracy_total = accuracy_total;
  if (st->res_pass_num == 1) {
    st->res_len_sd = 0.0;
    st->res_accuracy_sd = 0.0;
  } else {
    variance = 0.0;
    for (i = 0; i <= c->len_max; i++)
      if (c->freq_len[i] > 0) variance += pow((st->res_len_mean - i), 2) * c->freq_len[i];
    st->res_len_sd = sqrt(variance / st->res_pass_num);
    variance = 0.0;
    for (i = 0; i <= 100000; i++)
      if (c->freq_acc[i] > 0) variance += pow((st->res_accuracy_mean - i * 0.00001), 2) * c->freq_acc[i];
    st->res_accuracy_sd = sqrt(variance / st->res_pass_num);
  }
  st->res_depth = depth_len > 0 ? (double)st->res_len_total / depth_len / c->pass_num : 0.0;
  st->res_sub_rate = (double)st->res_sub_num / st->res_len_total;
  st->res_ins_rate = (double)st->res_ins_num / st->res_len_total;
  st->res_del_rate = (double)st->res_del_num / st->res_len_total;
}

/* ref: simulate_by_qshmm :2172-2410, simulate_by_errhmm :3791-4105, with init_sim_res :1437
 * and the quota from main :705. */
int orc_simulate_wgs(orc_ctx *c, double depth) {
  long long len_quota, len_total = 0;
  double accuracy_total = 0.0;
  rng_t *r = &c->rng;
  orc_stats_t *st = &c->st;

  if (!c->tables_built) return fail(c, "tables not built");
  if (!c->seq) return fail(c, "no sequence");
  if (r->mode == RNG_PHILOX) r->key[1] = (uint32_t)c->seq_num;
  r->scheme = (c->method == ORC_METHOD_ERR) ? 0 : 1;
  c->set_mode = 0;
  begin_stats(c);
  len_quota = (long long)(depth * c->glen);

  while (len_total < len_quota) {
    long index, wlen, offset, len0 = 0;
    int acc;
    char strand;
    int64_t start_draw = r->cur;

    d_plan_begin(r, (uint32_t)(st->res_num + 1));
    index = d_plan_len(r, (uint32_t)c->len_rand_value) + 1;
    wlen = c->prob2len[index];
    if (len_total + wlen > len_quota) {
      wlen = (long)(len_quota - len_total);
      if (wlen < c->len_min) wlen = c->len_min;
    }
    index = d_plan_acc(r, (uint32_t)c->accuracy_rand_value) + 1;
    acc = (int)c->prob2acc[index];
    if (wlen >= c->glen) {
      offset = 0;
      wlen = (long)c->glen;
    } else {
      offset = (long)d_plan_off(r, (uint64_t)(c->glen - wlen + 1));
    }
    st->res_num++;
    strand = (st->res_num % 2 == 1) ? '+' : '-';
    load_window(c, offset, wlen, strand);
    run_read_passes(c, offset, wlen, strand, acc, start_draw, &len0, &accuracy_total);
    len_total += len0;
    if (r->exhausted) return fail(c, "draw log exhausted");
  }
  finish_stats(c, accuracy_total, c->glen);
  return 0;
}

/* ------------------------------------------------------------------ transcript / template strategies */

/* ref: the "sequencing start pos distribution" table :2504-2528 (= :4193-4217).  ends[rank*21 + j-1] is the
 * cumulative table position of outcome j (start fraction (j-1)*5 %), mod[rank] the row modulus. */
static void build_ssp(int rank_max, long *ends, long *mod) {
  long i, j;
  for (i = 1; i <= rank_max; i++) {
    double sum = 0, value = (double)1 / i, ssp_prob_total = 0.0;
    long end_wk = 0;
    for (j = 1; j <= 21; j++) sum += value / pow(j, (1 + value));
    for (j = 1; j <= 21; j++) ends[i * 21 + j - 1] = -1;
    for (j = 1; j <= 21; j++) {
      ssp_prob_total += (value / pow(j, (1 + value))) / sum;
      end_wk = trunc_int(ssp_prob_total * 1000 + 0.5);
      if (end_wk > 1000) end_wk = 1000;
      ends[i * 21 + j - 1] = end_wk;
      if (end_wk >= 1000) break;
    }
    mod[i] = end_wk;
  }
}

static long ssp_lookup(const long *ends, int rank, long index /* 1-based */) {
  int j;
  for (j = 1; j <= 21; j++) {
    long e = ends[rank * 21 + j - 1];
    if (e < 0) break;
    if (index <= e) return (j - 1) * 5;
  }
  return 100; /* unreachable: index <= modulus = last end */
}

/* ref: simulate_by_{qshmm,errhmm}_trans :2419 / :4114 (strategy 1) and _templ :3055 / :4807 (strategy 2).
 * The file parsing of the reference (get_transcript_inf :1075, get_templ_inf :1366 and the fgets loops) is the
 * caller's: sequences arrive concatenated in `bases` with start[n+1]; names in `ids` with id_start[n+1]. */
int orc_simulate_set(orc_ctx *c, int strategy, int64_t n, const char *bases, const int64_t *start,
                     const int32_t *plus_exp, const int32_t *minus_exp, const char *ids, const int32_t *id_start) {
  rng_t *r = &c->rng;
  orc_stats_t *st = &c->st;
  double accuracy_total = 0.0;
  int64_t t, max_len = 0;
  long *ssp_ends = NULL, *ssp_mod = NULL;
  int rank_max = 0, i;
  /* which loops upper-case from index 0 (see hp_scan_w) */
  const int upper_from = (strategy == 1 && c->method == ORC_METHOD_QS) ? 0 : 1;
  char name[256];

  if (!c->tables_built) return fail(c, "tables not built");
  if (strategy != 1 && strategy != 2) return fail(c, "strategy must be 1 (transcript) or 2 (template)");
  if (r->mode == RNG_PHILOX) r->key[1] = 0;
  r->scheme = (c->method == ORC_METHOD_ERR) ? 0 : 1;
  c->set_mode = strategy;
  for (t = 0; t < n; t++)
    if (start[t + 1] - start[t] > max_len) max_len = start[t + 1] - start[t];
  if (strategy == 1) {
    rank_max = (int)ceil((float)max_len / 1000); /* ref: :1137 */
    ssp_ends = (long *)malloc((size_t)(rank_max + 1) * 21 * sizeof(long));
    ssp_mod = (long *)calloc((size_t)rank_max + 1, sizeof(long));
    build_ssp(rank_max, ssp_ends, ssp_mod);
  }
  begin_stats(c);

  /* --hp-del-bias != 1: frequency prepass over the whole file (ref: :2671-2746 trans, :3244-3310 templ).
   * hpfreq[0..10] are zeroed, [11] (which aliases hp_del_bias[0]) is not. */
  if (c->hp_del_bias_opt == 1) {
    for (i = 1; i <= 10; i++) c->bias[i] = 1;
  } else {
    for (i = 0; i <= 10; i++) c->hpfreq[i] = 0;
    for (t = 0; t < n; t++) {
      int64_t len = start[t + 1] - start[t];
      long weight = strategy == 1 ? (long)plus_exp[t] + (long)minus_exp[t] : 1;
      char *tmp = (char *)malloc((size_t)len + 1);
      memcpy(tmp, bases + start[t], (size_t)len);
      tmp[len] = 0;
      hp_scan_w(c, tmp, len, NULL, upper_from, weight, 1);
      free(tmp);
    }
    orc_finish_bias(c);
  }
  refresh_bias0(c);

  for (t = 0; t < n; t++) {
    int64_t len = start[t + 1] - start[t];
    long read_num = strategy == 1 ? (long)plus_exp[t] + (long)minus_exp[t] : 1, k;
    int idn = id_start[t + 1] - id_start[t];
    if (idn > 128) idn = 128; /* TRANS_ID_LEN_MAX / REF_ID_LEN_MAX */
    memcpy(name, ids + id_start[t], (size_t)idn);
    name[idn] = 0;
    c->rec_name = name;
    c->rec_name_width = strategy == 1 ? idn : 3; /* ref: :2968 vs :3486 */
    /* the sequence becomes the "genome" of its reads */
    free(c->seq);
    free(c->hp_alloc);
    c->seq = (char *)malloc((size_t)len + 1);
    memcpy(c->seq, bases + start[t], (size_t)len);
    c->seq[len] = 0;
    c->hp_alloc = (int16_t *)calloc((size_t)len + 2, sizeof(int16_t));
    c->hp = c->hp_alloc + 1;
    c->glen = len;
    hp_scan_w(c, c->seq, len, c->hp, upper_from, 1, 0);

    for (k = 1; k <= read_num; k++) {
      long index, wlen, offset, len0 = 0;
      int acc;
      char strand;
      int64_t start_draw = r->cur;
      d_plan_begin(r, (uint32_t)(st->res_num + 1));
      if (strategy == 1) { /* ref: :2842-2866 */
        int rank;
        long ssp;
        double value;
        index = d_plan_len(r, (uint32_t)c->len_rand_value) + 1;
        wlen = c->prob2len[index];
        index = d_plan_acc(r, (uint32_t)c->accuracy_rand_value) + 1;
        acc = (int)c->prob2acc[index];
        rank = (int)ceil((double)len / 1000);
        index = (long)d_plan_off(r, (uint64_t)ssp_mod[rank]) + 1;
        ssp = ssp_lookup(ssp_ends, rank, index);
        value = ssp == 0 ? 0.0 : ((double)ssp - 2.5) / 100;
        offset = trunc_int((double)len * value + 0.5);
        if (offset + wlen > len) wlen = (long)len - offset;
        strand = (k <= plus_exp[t]) ? '+' : '-';
      } else { /* ref: :3359-3364 */
        index = d_plan_acc(r, (uint32_t)c->accuracy_rand_value) + 1;
        acc = (int)c->prob2acc[index];
        offset = 0;
        wlen = (long)len;
        strand = '+';
      }
      if (wlen < 1) {
        free(ssp_ends); free(ssp_mod);
        return fail(c, "empty read window: the reference divides by zero here (sequences must be longer than 20 bases)");
      }
      st->res_num++;
      load_window(c, offset, wlen, strand);
      run_read_passes(c, offset, wlen, strand, acc, start_draw, &len0, &accuracy_total);
      if (r->exhausted) {
        free(ssp_ends); free(ssp_mod);
        return fail(c, "draw log exhausted");
      }
      /* Reference quirk (simulate_by_errhmm_trans only): the verbatim copy of an accuracy-100 read is
       * `for (i=0; i<mut.len; i++)` (:4532) with the SAME i that counts the transcript's reads (:4487), so after
       * such a read the count continues from mut.len + 1 — usually past read_num: the transcript's remaining reads
       * are never simulated.  Reproduced for the reference's own draw stream; PHILOX mode (the engine's numbering of
       * reads is static) simulates every read. */
      if (strategy == 1 && c->method == ORC_METHOD_ERR && acc == 100 && r->mode != RNG_PHILOX) k = wlen;
    }
  }
  free(ssp_ends);
  free(ssp_mod);
  c->rec_name = NULL;
  finish_stats(c, accuracy_total, 0);
  return 0;
}

int64_t orc_get_ssp(int rank_max, int32_t *ends /* [(rank_max+1)*21] */, int32_t *mod /* [rank_max+1] */) {
  long *e = (long *)malloc((size_t)(rank_max + 1) * 21 * sizeof(long));
  long *m = (long *)calloc((size_t)rank_max + 1, sizeof(long));
  int64_t i;
  build_ssp(rank_max, e, m);
  for (i = 21; i < (int64_t)(rank_max + 1) * 21; i++) ends[i] = (int32_t)e[i];
  for (i = 1; i <= rank_max; i++) mod[i] = (int32_t)m[i];
  free(e);
  free(m);
  return rank_max;
}

/* ------------------------------------------------------------------ --method sample (engine: k_sim_sample, sample_plan.hpp) */

/* ref: simulate_by_sample inner loop :1775-1833.  The quality string of the sampled read gives the quality of
 * every position; `len` bounds BOTH the window and the read (mut.len, :1756-1763), and no deletion is drawn once
 * either is used up. */
static void sample_pass(orc_ctx *c, rng_t *r, const char *q, long len, pass_out_t *po) {
  long ref_offset = 0, read_offset = 0, maf_offset = 0, index, qv = 0, rand_value;
  char nt;
  po->nsub = po->nins = po->ndel = 0;
  while ((ref_offset < len) && (read_offset < len)) {
    d_begin(r, 0, (uint32_t)read_offset);
    nt = c->w_seq[ref_offset];
    qv = (int)q[read_offset] - 33;
    c->qual[read_offset] = q[read_offset];
    rand_value = d_qs_err(r);
    if (d_lt(r, (uint32_t)rand_value, (double)c->sub_thre[qv])) {
      po->nsub++;
      c->read_seq[read_offset] = substitute(r, nt);
      c->maf_ref[maf_offset] = nt;
      ref_offset++;
    } else if (d_lt(r, (uint32_t)rand_value, (double)c->ins_thre[qv])) {
      po->nins++;
      index = d_choice8(r);
      c->read_seq[read_offset] = (index >= 4) ? nt : NT4[index];
      c->maf_ref[maf_offset] = '-';
    } else {
      c->read_seq[read_offset] = nt;
      c->maf_ref[maf_offset] = nt;
      ref_offset++;
    }
    c->maf_seq[maf_offset] = c->read_seq[read_offset];
    maf_offset++;
    read_offset++;
    {
      uint32_t j = 0;
      while ((ref_offset < len) && (read_offset < len)) {
        int hp = c->w_hp[ref_offset - 1];
        rand_value = d_del(r, j++);
        if (d_lt(r, (uint32_t)rand_value, c->del_thre[qv] * c->bias[hp])) {
          po->ndel++;
          c->maf_seq[maf_offset] = '-';
          c->maf_ref[maf_offset] = c->w_seq[ref_offset];
          maf_offset++;
          ref_offset++;
        } else {
          break;
        }
      }
    }
  }
  po->rlen = read_offset;
  po->ncol = maf_offset;
}

/* the draw that starts pass p over the pool (ref: sample_value = rand() % sample.num_filtered, :1734) */
static uint32_t d_pool_pass(rng_t *r, uint32_t p, uint32_t n) {
  if (r->mode == RNG_PHILOX) {
    uint32_t ctr[4] = {p, 0u, 0u, 3u}, w[4];
    orc_philox4x32_10(ctr, r->key, w);
    return mulhi32(w[0], n);
  }
  return stream_next(r) % n;
}

/* ref: simulate_by_sample :1694-1949.  The pool is what get_sample_inf (:1214-1275) leaves in fp_filtered: the
 * quality strings that pass the length and accuracy filters, in file order (quals / qstart[n+1]).
 * Quirk kept: the quality buffer is cut at the end of every read (mut.qc[read_offset] = 0, :1835) and measured again
 * for the next copy of the same pool entry (mut.len = strlen(mut.qc), :1756): copy i+1 is as long as copy i's READ. */
int orc_simulate_sample(orc_ctx *c, double depth, int64_t n, const char *quals, const int64_t *qstart) {
  rng_t *r = &c->rng;
  orc_stats_t *st = &c->st;
  long long len_quota, len_total = 0, pool_total;
  long sample_num, sample_interval, sample_value, sample_residue, num, i;
  double accuracy_total = 0.0;
  uint32_t pool_pass = 0;
  int64_t j;

  if (!c->seq) return fail(c, "no sequence");
  if (n < 2) return fail(c, "the reference divides by zero with a pool of fewer than 2 reads (:1723)");
  if (r->mode == RNG_PHILOX) r->key[1] = (uint32_t)c->seq_num;
  r->scheme = 1;
  c->set_mode = 0;
  begin_stats(c);
  pool_total = qstart[n];
  len_quota = (long long)(depth * c->glen);
  sample_num = (long)(len_quota / pool_total);
  sample_residue = (long)(len_quota % pool_total);
  if (sample_residue == 0) {
    sample_interval = 1;
  } else {
    sample_interval = trunc_int((double)(pool_total / sample_residue) * 2 + 0.5);
    if (sample_interval > (long)(n * 0.5)) sample_interval = (long)(n * 0.5);
  }

  while (len_total < len_quota) {
    int64_t pass_draw = r->cur;
    sample_value = (long)d_pool_pass(r, pool_pass++, (uint32_t)n);
    for (j = 0; j < n; j++) {
      const char *q = quals + qstart[j];
      long curlen = (long)(qstart[j + 1] - qstart[j]);  /* strlen(mut.qc) after fgets + trim */
      if (len_total >= len_quota) break;
      num = (sample_value % sample_interval == 0) ? sample_num + 1 : sample_num;
      sample_value++;
      for (i = 0; i < num; i++) {
        long len, offset, k;
        char strand;
        pass_out_t po;
        orc_readinfo_t ri;
        double prob = 0.0, value;
        int64_t start_draw = pass_draw >= 0 ? pass_draw : r->cur;
        if (len_total >= len_quota) break;
        pass_draw = -1;
        len = curlen;
        d_plan_begin(r, (uint32_t)(st->res_num + 1));
        if (len >= c->glen) {
          offset = 0;
          len = (long)c->glen;
        } else {
          offset = (long)d_plan_off(r, (uint64_t)(c->glen - len + 1));
        }
        st->res_num++;
        strand = (st->res_num % 2 == 1) ? '+' : '-';
        load_window(c, offset, len, strand);
        sample_pass(c, r, q, len, &po);
        curlen = po.rlen; /* the quirk: the buffer now ends where this read ended */
        if (strand == '-') {
          revcomp_n(c->maf_seq, po.ncol);
          revcomp_n(c->maf_ref, po.ncol);
        }
        st->res_sub_num += po.nsub;
        st->res_ins_num += po.nins;
        st->res_del_num += po.ndel;
        st->res_len_total += po.rlen;
        len_total += po.rlen;
        if (po.rlen < c->freq_len_n) c->freq_len[po.rlen]++;
        if (po.rlen > st->res_len_max) st->res_len_max = po.rlen;
        if (po.rlen < st->res_len_min) st->res_len_min = po.rlen;
        if (r->mode == RNG_PHILOX) prob = qs_prob_sum_fixed(c, po.rlen);
        else for (k = 0; k < po.rlen; k++) prob += c->qc_prob[(int)c->qual[k] - 33];
        value = 1.0 - (prob / po.rlen);
        accuracy_total += value;
        {
          long acc_wk = trunc_int(value * 100000 + 0.5);
          if (acc_wk >= 0 && acc_wk <= 100000) c->freq_acc[acc_wk]++;
        }
        /* the MAF reference row covers what was consumed (mut.seq_right = offset + ref_offset, :1847) */
        emit_records(c, (long)st->res_num, 0, offset, po.ncol - po.nins, strand, po.rlen, po.ncol);
        ri.read_id = st->res_num; ri.pass = 0; ri.acc = 0; ri.offset = offset; ri.wlen = len; ri.rlen = po.rlen;
        ri.ncol = po.ncol; ri.strand = strand; ri.nsub = (int32_t)po.nsub; ri.nins = (int32_t)po.nins;
        ri.ndel = (int32_t)po.ndel; ri.draw_start = start_draw; ri.accuracy = value;
        push_info(c, &ri);
        if (r->exhausted) return fail(c, "draw log exhausted");
      }
    }
    sample_num = 0;
  }
  {
    int keep = c->pass_num;
    c->pass_num = 1; /* ref: means over res_num (:1923-1946) */
    finish_stats(c, accuracy_total, c->glen);
    c->pass_num = keep;
  }
  return 0;
}

/* ------------------------------------------------------------------ getters */
const char *orc_out_reads(orc_ctx *c, int64_t *n) { *n = c->out_reads.n; return c->out_reads.p; }
const char *orc_out_maf(orc_ctx *c, int64_t *n) { *n = c->out_maf.n; return c->out_maf.p; }
void orc_get_stats(orc_ctx *c, orc_stats_t *st) { *st = c->st; }
const orc_readinfo_t *orc_get_readinfo(orc_ctx *c, int64_t *n) { *n = c->info_n; return c->info; }
const int32_t *orc_get_draw_log(orc_ctx *c, int64_t *n) { *n = c->rng.cur; return c->rng.rec; }
int64_t orc_draws_consumed(orc_ctx *c) { return c->rng.cur; }
const int64_t *orc_get_freq_len(orc_ctx *c, int64_t *n) { *n = c->freq_len_n; return c->freq_len; }
const int64_t *orc_get_freq_accuracy(orc_ctx *c, int64_t *n) { *n = 100001; return c->freq_acc; }

int64_t orc_get_table(orc_ctx *c, int which, int acc, int state, int32_t *out, int64_t cap) {
  int64_t n = 0, k;
  int err = (c->method == ORC_METHOD_ERR);
  switch (which) {
    case 0: n = c->len_rand_value; for (k = 0; k < n && k < cap; k++) out[k] = (int32_t)c->prob2len[k + 1]; break;
    case 1: n = c->accuracy_rand_value; for (k = 0; k < n && k < cap; k++) out[k] = (int32_t)c->prob2acc[k + 1]; break;
    case 2: n = c->mod_init[acc];
      for (k = 0; k < n && k < cap; k++) out[k] = err ? c->er_init[acc][k + 1] : c->qs_init[acc][k + 1];
      break;
    case 3: n = c->mod_emis[acc][state];
      for (k = 0; k < n && k < cap; k++) out[k] = err ? c->er_emis[acc][state][k + 1] : c->qs_emis[acc][state][k + 1];
      break;
    case 4: n = c->mod_tran[acc][state];
      for (k = 0; k < n && k < cap; k++) out[k] = err ? c->er_tran[acc][state][k + 1] : c->qs_tran[acc][state][k + 1];
      break;
    case 5: n = c->mod_freq[acc]; for (k = 0; k < n && k < cap; k++) out[k] = c->qs_freq[acc][k + 1]; break;
    default: break;
  }
  return n;
}
int orc_get_emis2del(orc_ctx *c, int acc, int state) { return (int)c->er_del[acc][state]; }
void orc_get_thresholds(orc_ctx *c, int64_t sub[94], int64_t ins[94], int64_t del[94]) {
  int i;
  for (i = 0; i < 94; i++) { sub[i] = c->sub_thre[i]; ins[i] = c->ins_thre[i]; del[i] = c->del_thre[i]; }
}
int orc_model_exists(orc_ctx *c, int acc) { return c->exist[acc]; }
void orc_model_range(orc_ctx *c, int *acc_min, int *acc_max, int *tab_acc_lo, int *tab_acc_hi) {
  *acc_min = c->acc_min; *acc_max = c->acc_max; *tab_acc_lo = (int)c->tab_acc_lo; *tab_acc_hi = (int)c->tab_acc_hi;
}
