/* TEST INFRASTRUCTURE — CPU oracle for the PBSIM3 read-generation hot path.
 *
 * This is a plain-C restatement of the algorithm in /root/reference/src/pbsim.cpp
 * (cited function by function in pbsim_oracle.c).  It exists ONLY to check the CUDA
 * engine: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it;
 * nothing under pbsim_b200/ links, imports or calls it.
 *
 * Parity pin: tests/test_oracle_vs_reference.py compares this oracle, driven by the
 * glibc rand() restatement (glibc_rand.c), byte for byte with outputs of the
 * unmodified reference binary (oracle/_ref/pbsim, built by `make -C oracle ref`)
 * captured in tests/golden/ by oracle/make_golden.py.
 */
#ifndef PBSIM_ORACLE_H
#define PBSIM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ACC_MAX 100
#define ORC_STATE_MAX 50
#define ORC_NQV 94
#define ORC_METHOD_QS 1
#define ORC_METHOD_ERR 2

typedef struct {
  int32_t r[31];
  int f, b;
} orc_glibc_rand_t;

void orc_glibc_srand(orc_glibc_rand_t *g, uint32_t seed);
int32_t orc_glibc_rand(orc_glibc_rand_t *g);
void orc_glibc_rand_fill(uint32_t seed, int64_t n, int32_t *out);

/* Philox KAT helper for the tests */
void orc_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

typedef struct orc_ctx orc_ctx;

typedef struct {
  int64_t res_num;
  int64_t res_pass_num;
  int64_t res_len_total;
  int64_t res_len_min, res_len_max;
  int64_t res_sub_num, res_ins_num, res_del_num;
  double res_depth;
  double res_len_mean, res_len_sd;
  double res_accuracy_mean, res_accuracy_sd;
  double res_sub_rate, res_ins_rate, res_del_rate;
  double accuracy_total;
} orc_stats_t;

/* one record per simulated (read, pass) */
typedef struct {
  int64_t read_id;     /* sim.res_num, 1-based */
  int32_t pass;
  int32_t acc;
  int64_t offset;      /* mut.offset */
  int64_t wlen;        /* mut.len (window) */
  int64_t rlen;        /* emitted read length */
  int64_t ncol;        /* alignment columns (MAF row length) */
  int32_t strand;      /* '+' or '-' */
  int32_t nsub, nins, ndel;
  int64_t draw_start;  /* first draw index of this subread (stream modes) */
  double accuracy;     /* per-read accuracy value */
} orc_readinfo_t;

orc_ctx *orc_new(void);
void orc_free(orc_ctx *c);
const char *orc_error(orc_ctx *c);

/* parameters: the subset of sim_t (pbsim.cpp:51-77) the hot path reads.
 * accuracy_mean must already be truncated as set_sim_param does (pbsim.cpp:1660). */
int orc_set_params(orc_ctx *c, int method, int pass_num, double accuracy_mean,
                   long len_min, long len_max, double len_mean, double len_sd,
                   long sub_ratio, long ins_ratio, long del_ratio,
                   double hp_del_bias, const char *id_prefix);
int orc_load_model(orc_ctx *c, const char *path);
int orc_build_tables(orc_ctx *c);

/* genome: set_sequence = get_genome_seq (pbsim.cpp:997-1068) for sequence seq_num.
 * prepass_sequence/finish_bias = the --hp-del-bias != 1 prepass (pbsim.cpp:673-697). */
int orc_prepass_sequence(orc_ctx *c, const char *seq, int64_t len);
int orc_finish_bias(orc_ctx *c);
int orc_set_sequence(orc_ctx *c, const char *seq, int64_t len, int seq_num);
void orc_get_bias(orc_ctx *c, double bias[12]);
const int16_t *orc_get_hp(orc_ctx *c, int64_t *n);
const char *orc_get_seq(orc_ctx *c, int64_t *n);

/* draw sources */
int orc_rng_glibc(orc_ctx *c, uint32_t seed);                     /* srand(seed) */
int orc_rng_replay(orc_ctx *c, const int32_t *log, int64_t n);    /* consume a draw log */
int orc_rng_philox(orc_ctx *c, uint32_t seed);                    /* engine PHILOX addressing */

/* simulate_by_qshmm / simulate_by_errhmm for the current sequence (pbsim.cpp:1955, :3594).
 * Output is appended to the context's buffers; call orc_reset_outputs between sequences. */
int orc_simulate_wgs(orc_ctx *c, double depth);
void orc_reset_outputs(orc_ctx *c);

/* simulate_by_{qshmm,errhmm}_trans (strategy 1, pbsim.cpp:2419 / :4114) and _templ (strategy 2, :3055 / :4807)
 * over n sequences given concatenated (bases/start[n+1], ids/id_start[n+1]); plus_exp / minus_exp are the
 * expression counts of the transcript table (ignored for templates: one '+' read each). */
int orc_simulate_set(orc_ctx *c, int strategy, int64_t n, const char *bases, const int64_t *start,
                     const int32_t *plus_exp, const int32_t *minus_exp, const char *ids, const int32_t *id_start);
/* simulate_by_sample (pbsim.cpp:1694) for the current sequence; the pool holds the quality strings that passed
 * get_sample_inf's filters (:1214-1275), in file order.  orc_set_params is enough (no model, no tables).
 * The CUDA engine does not implement this method yet: the restatement pins the specification for it. */
int orc_simulate_sample(orc_ctx *c, double depth, int64_t n, const char *quals, const int64_t *qstart);
/* start-position table (pbsim.cpp:2504-2528) for KATs: ends[rank*21 + j-1], mod[rank], rank = 1..rank_max */
int64_t orc_get_ssp(int rank_max, int32_t *ends, int32_t *mod);

const char *orc_out_reads(orc_ctx *c, int64_t *n); /* FASTQ (pass_num==1) or SAM records */
const char *orc_out_maf(orc_ctx *c, int64_t *n);
void orc_get_stats(orc_ctx *c, orc_stats_t *st);
const orc_readinfo_t *orc_get_readinfo(orc_ctx *c, int64_t *n);
const int32_t *orc_get_draw_log(orc_ctx *c, int64_t *n);   /* draws consumed so far (glibc mode) */
int64_t orc_draws_consumed(orc_ctx *c);
const int64_t *orc_get_freq_len(orc_ctx *c, int64_t *n);
const int64_t *orc_get_freq_accuracy(orc_ctx *c, int64_t *n);

/* quantised tables, for KATs against the product's host table builder.
 * which: 0 prob2len, 1 prob2accuracy, 2 init2state[acc], 3 emis[acc][state], 4 tran[acc][state],
 *        5 freq2qc[acc] ; returns number of valid entries (= the row modulus) and copies
 *        min(cap, n) int32 values (0-based: out[k] is the reference's table[k+1]). */
int64_t orc_get_table(orc_ctx *c, int which, int acc, int state, int32_t *out, int64_t cap);
int orc_get_emis2del(orc_ctx *c, int acc, int state);
void orc_get_thresholds(orc_ctx *c, int64_t sub[94], int64_t ins[94], int64_t del[94]);
int orc_model_exists(orc_ctx *c, int acc);
void orc_model_range(orc_ctx *c, int *acc_min, int *acc_max, int *tab_acc_lo, int *tab_acc_hi);

#ifdef __cplusplus
}
#endif
#endif
