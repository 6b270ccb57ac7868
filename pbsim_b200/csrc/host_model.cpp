// Host front end of libpbsim_cuda: model file parser and quantised-table builder.
//
// Produces the `pbsim_model` the engine uploads.  Behaviour follows the reference
// (yukiteruono/pbsim3, src/pbsim.cpp) line by line where results depend on it:
//   model text format and storage aliasing      set_qshmm :5570-5634, set_errhmm :5640-5714
//   Phred table, uniform-QV fallback            main :546-578
//   sub/ins/del thresholds                      set_mut :5474-5479
//   length / accuracy samplers                  simulate_by_qshmm :1991-2064 (= :3633-3706)
//   HMM lookup tables                           :2066-2170 (qshmm), :3708-3789 (errhmm)
// No GPU code here; compiled into libpbsim_cuda.so so that host drivers get one library.
#include "../../include/pbsim_cuda.h"
#include "gz_host.hpp"

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int kAccCells = PBSIM_NACC;  // 0..100
constexpr int kStateCells = 51;        // STATE_MAX + 1 (pbsim.cpp:43)
constexpr int kStateMax = 50;
constexpr size_t kLineBytes = 10240;   // BUF_SIZE (pbsim.cpp:20): longer lines are split by fgets

// The reference writes `int(expr)`; on x86-64 that is cvttsd2si, which returns INT_MIN for NaN
// and out-of-range values (SURVEY App. B-12 relies on this for the "not appropriate" errors).
inline long cxx_int(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return static_cast<long>(INT_MIN);
  return static_cast<long>(static_cast<int>(v));
}

// Raw probabilities, stored flat with the reference's array geometry so that state numbers
// beyond STATE_MAX land in the same neighbouring cells as in the reference
// (QSHMM-ONT-HQ has 56 states).
struct RawModel {
  int ep_cols = 94;
  std::vector<double> ip, ep, tp;
  int exist[kAccCells] = {0};
  int state_max[kAccCells] = {0};
  int acc_min = 100, acc_max = 0;

  void reset(int cols) {
    ep_cols = cols;
    ip.assign(static_cast<size_t>(kAccCells) * kStateCells, 0.0);
    ep.assign(static_cast<size_t>(kAccCells) * kStateCells * cols, 0.0);
    tp.assign(static_cast<size_t>(kAccCells) * kStateCells * kStateCells, 0.0);
  }
  double IP(int a, int s) const { return ip[static_cast<size_t>(a) * kStateCells + s]; }
  double EP(int a, int s, int k) const { return ep[(static_cast<size_t>(a) * kStateCells + s) * ep_cols + k]; }
  double TP(int a, int s, int k) const { return tp[(static_cast<size_t>(a) * kStateCells + s) * kStateCells + k]; }
};

bool store(std::vector<double> &arr, long long idx, double v) {
  if (idx < 0 || idx >= static_cast<long long>(arr.size())) return false;
  arr[static_cast<size_t>(idx)] = v;
  return true;
}

// "<acc> IP <state> <p>" | "<acc> EP <state> <p_0 ...>" | "<acc> TP <state> <p_1 ...>"
int parse_model(const char *path, int method, RawModel &raw, const char **err) {
  FILE *fp = std::fopen(path, "r");
  if (!fp) {
    *err = "ERROR: Cannot open file (model)";
    return PBSIM_E_IO;
  }
  raw.reset(method == PBSIM_METHOD_ERRHMM ? 4 : 94);
  std::vector<char> buf(kLineBytes);
  int rc = 0;
  while (std::fgets(buf.data(), static_cast<int>(kLineBytes), fp)) {
    size_t n = std::strlen(buf.data());
    if (n && buf[n - 1] == '\n') buf[n - 1] = '\0';
    char *save = nullptr;
    char *field = strtok_r(buf.data(), " ", &save);
    if (!field) continue;
    const int acc = std::atoi(field);
    if (acc < 0 || acc >= kAccCells) { rc = PBSIM_E_IO; *err = "model: accuracy out of range"; break; }
    raw.exist[acc] = 1;
    if (acc < raw.acc_min) raw.acc_min = acc;
    if (acc > raw.acc_max) raw.acc_max = acc;
    char *kind = strtok_r(nullptr, " ", &save);
    char *st = kind ? strtok_r(nullptr, " ", &save) : nullptr;
    if (!kind || !st) continue;
    const int state = std::atoi(st);
    const long long row = static_cast<long long>(acc) * kStateCells + state;
    bool ok = true;
    if (std::strcmp(kind, "IP") == 0) {
      char *v = strtok_r(nullptr, " ", &save);
      ok = v && store(raw.ip, row, std::atof(v));
      raw.state_max[acc] = state;
    } else if (std::strcmp(kind, "EP") == 0) {
      int col = 0;  // emission columns are 0-based (:5614)
      for (char *v = strtok_r(nullptr, " ", &save); v && ok; v = strtok_r(nullptr, " ", &save))
        ok = store(raw.ep, row * raw.ep_cols + col++, std::atof(v));
    } else if (std::strcmp(kind, "TP") == 0) {
      int col = 0;  // transition columns are 1-based (:5625)
      for (char *v = strtok_r(nullptr, " ", &save); v && ok; v = strtok_r(nullptr, " ", &save))
        ok = store(raw.tp, row * kStateCells + ++col, std::atof(v));
    }
    if (!ok) { rc = PBSIM_E_IO; *err = "model: entry outside the reference's arrays"; break; }
  }
  std::fclose(fp);
  return rc;
}

// One quantised CDF row.  `carry` plays the role of the reference's function-scope end_wk: it is
// only assigned when an outcome with non-zero probability is visited, so an empty row inherits
// the previous row's modulus exactly as the reference does.
struct Quantiser {
  long carry = 0;

  template <class ProbFn>
  long fill(uint8_t *dst, long resolution, int first, int last, ProbFn prob, bool skip_nonpositive) {
    long lo = 1;
    double cum = 0.0;
    for (int k = first; k <= last; ++k) {
      const double p = prob(k);
      if (skip_nonpositive ? (p <= 0) : (p == 0)) continue;
      cum += p;
      long hi = cxx_int(cum * resolution + 0.5);
      if (hi > resolution) hi = resolution;
      for (long t = lo; t <= hi; ++t) dst[t - 1] = static_cast<uint8_t>(k);
      carry = hi;
      if (hi >= resolution) break;
      lo = hi + 1;
    }
    return carry;
  }
};

}  // namespace

struct pbsim_host_model {
  pbsim_model view;
  RawModel raw;
  std::vector<int32_t> prob2len;
  std::vector<uint8_t> prob2acc;
  struct Row {
    std::vector<uint8_t> init, tran, emis, freq;
    std::vector<int32_t> tran_mod, emis_mod, emis_del;
  };
  Row rows[kAccCells];
};

namespace {

int build_samplers(pbsim_host_model &m, const pbsim_host_params &p, const char **err) {
  pbsim_model &v = m.view;
  // length: Gamma(kappa = mean^2/sd^2, theta = sd^2/mean) pdf summed over integer lengths
  const double variance = std::pow(p.len_sd, 2);
  const double kappa = std::pow(p.len_mean, 2) / variance;
  const double theta = variance / p.len_mean;
  const double gam = std::tgamma(kappa);
  std::vector<long> tab(100001, 0);
  long modulus = 0;
  if (p.len_sd == 0.0) {
    tab[1] = cxx_int(p.len_mean + 0.5);
    modulus = 1;
  } else {
    long lo = 1;
    double cum = 0.0;
    for (long len = p.len_min; len <= p.len_max; ++len) {
      cum += std::pow(static_cast<double>(len), kappa - 1) * std::exp(static_cast<double>(-1 * len) / theta) /
             std::pow(theta, kappa) / gam;
      long hi = cxx_int(cum * 100000 + 0.5);
      if (hi > 100000) hi = 100000;
      for (long t = lo; t <= hi; ++t) tab[t] = len;
      modulus = hi;
      if (hi >= 100000) break;
      lo = hi + 1;
    }
  }
  if (modulus < 1) {
    // the reference only checks this when pass_num == 1 (:2022) and divides by zero otherwise
    *err = "ERROR: length parameters are not appropriate.";
    return PBSIM_E_PARAM;
  }
  m.prob2len.assign(tab.begin() + 1, tab.begin() + 1 + modulus);
  v.prob2len = m.prob2len.data();
  v.len_rand_value = static_cast<int32_t>(modulus);

  // accuracy: P(acc = i) proportional to exp(0.22 i) on [floor(0.75 m), min(100, floor(1.05 m))]
  const double mean = p.accuracy_mean * 100;
  long hi_acc = static_cast<long>(std::floor(mean * 1.05));
  const long lo_acc = static_cast<long>(std::floor(mean * 0.75));
  if (hi_acc > 100) hi_acc = 100;
  double total = 0.0;
  for (long i = lo_acc; i <= hi_acc; ++i) total += std::exp(0.22 * i);
  std::vector<uint8_t> atab(100001, 0);
  long lo = 1, amod = 0;
  double cum = 0.0;
  for (long i = lo_acc; i <= hi_acc; ++i) {
    cum += std::exp(0.22 * i) / total;
    long hi = cxx_int(cum * 100000 + 0.5);
    if (hi > 100000) hi = 100000;
    for (long t = lo; t <= hi; ++t) atab[t] = static_cast<uint8_t>(i);
    amod = hi;
    if (hi >= 100000) break;
    lo = hi + 1;
  }
  if (amod < 1 || lo_acc < 0) {
    *err = "ERROR: accuracy parameters are not appropriate.";
    return PBSIM_E_PARAM;
  }
  m.prob2acc.assign(atab.begin() + 1, atab.begin() + 1 + amod);
  v.prob2accuracy = m.prob2acc.data();
  v.accuracy_rand_value = static_cast<int32_t>(amod);
  v.acc_lo = static_cast<int32_t>(lo_acc);
  v.acc_hi = static_cast<int32_t>(hi_acc);
  return 0;
}

void build_thresholds(pbsim_model &v, const pbsim_host_params &p, double uni_ep[kAccCells][PBSIM_NQV]) {
  const long sum = p.sub_ratio + p.ins_ratio + p.del_ratio;
  const double sub_rate = static_cast<double>(p.sub_ratio) / sum;
  const double ins_rate = static_cast<double>(p.ins_ratio) / sum;
  const double del_rate = static_cast<double>(p.del_ratio) / sum;
  for (int q = 0; q < PBSIM_NQV; ++q) v.qc_prob[q] = std::pow(10, static_cast<double>(q) / -10);
  for (int q = 0; q < PBSIM_NQV; ++q) {
    const double e = v.qc_prob[q];
    v.sub_thre[q] = static_cast<int32_t>(cxx_int((e * sub_rate) * 1000000 + 0.5));
    v.ins_thre[q] = static_cast<int32_t>(cxx_int((e * (sub_rate + ins_rate)) * 1000000 + 0.5));
    v.del_thre[q] = static_cast<int32_t>(cxx_int((e * del_rate) / (1 + e * del_rate) * 1000000 + 0.5));
  }
  // accuracies the model does not cover emit a two-point mixture of adjacent QVs whose mean
  // error probability is 1 - acc/100
  for (int a = 0; a < kAccCells; ++a) {
    for (int q = 0; q < PBSIM_NQV; ++q) uni_ep[a][q] = 0;
    if (a == 100) {
      uni_ep[a][93] = 1.0;
      continue;
    }
    const double target = 1.0 - a / 100.0;
    for (int q = 0; q < PBSIM_NQV; ++q) {
      if (target == v.qc_prob[q]) {
        uni_ep[a][q] = 1.0;
        break;
      }
      if (target > v.qc_prob[q]) {
        const double w = (target - v.qc_prob[q]) / (v.qc_prob[q - 1] - v.qc_prob[q]);
        uni_ep[a][q - 1] = w;
        uni_ep[a][q] = 1 - w;
        break;
      }
    }
  }
}

void build_hmm_rows(pbsim_host_model &m, double uni_ep[kAccCells][PBSIM_NQV]) {
  pbsim_model &v = m.view;
  const RawModel &raw = m.raw;
  const bool err_model = (v.method == PBSIM_METHOD_ERRHMM);
  const int res = err_model ? 1000 : 100;
  Quantiser q;
  q.carry = v.accuracy_rand_value;  // end_wk still holds the accuracy sampler's last value (:2059)
  for (int a = 0; a < kAccCells; ++a) {
    pbsim_hmm_row &r = v.rows[a];
    std::memset(&r, 0, sizeof r);
    r.exists = raw.exist[a];
    r.resolution = res;
  }
  for (int a = v.acc_lo; a <= v.acc_hi; ++a) {
    if (a < 0 || a >= kAccCells) continue;
    pbsim_hmm_row &r = v.rows[a];
    pbsim_host_model::Row &s = m.rows[a];
    if (!raw.exist[a]) {
      if (err_model) continue;  // errhmm borrows the nearest modelled accuracy at run time (:3872-3926)
      s.freq.assign(1000, 0);
      r.freq_mod = static_cast<int32_t>(
          q.fill(s.freq.data(), 1000, 0, 93, [&](int k) { return uni_ep[a][k]; }, false));
      r.freq = s.freq.data();
      continue;
    }
    const int nst = err_model ? raw.state_max[a] : kStateMax;
    r.nstates = nst;
    const size_t cells = static_cast<size_t>(nst + 1) * res;
    s.init.assign(res, 0);
    s.tran.assign(cells, 0);
    s.emis.assign(cells, 0);
    s.tran_mod.assign(nst + 1, 0);
    s.emis_mod.assign(nst + 1, 0);
    r.init_mod = static_cast<int32_t>(q.fill(s.init.data(), res, 1, nst, [&](int k) { return raw.IP(a, k); }, false));
    if (err_model) {
      s.emis_del.assign(nst + 1, 0);
      for (int st = 1; st <= nst; ++st) {
        s.emis_del[st] = static_cast<int32_t>(cxx_int(raw.EP(a, st, 3) * 1000 + 0.5));
        s.emis_mod[st] = static_cast<int32_t>(
            q.fill(&s.emis[static_cast<size_t>(st) * res], res, 0, 2, [&](int k) { return raw.EP(a, st, k); }, true));
      }
    } else {
      for (int st = 1; st <= nst; ++st)
        s.emis_mod[st] = static_cast<int32_t>(
            q.fill(&s.emis[static_cast<size_t>(st) * res], res, 0, 93, [&](int k) { return raw.EP(a, st, k); }, false));
    }
    for (int st = 1; st <= nst; ++st)
      s.tran_mod[st] = static_cast<int32_t>(
          q.fill(&s.tran[static_cast<size_t>(st) * res], res, 1, kStateMax, [&](int k) { return raw.TP(a, st, k); }, false));
    r.init = s.init.data();
    r.tran = s.tran.data();
    r.emis = s.emis.data();
    r.tran_mod = s.tran_mod.data();
    r.emis_mod = s.emis_mod.data();
    r.emis_del = err_model ? s.emis_del.data() : nullptr;
  }
}

}  // namespace

extern "C" {

int pbsim_host_model_load(pbsim_host_model **out, const pbsim_host_params *p, const char *model_path,
                          const char **err) {
  static const char *dummy;
  if (!err) err = &dummy;
  *err = "";
  const bool sample = p && p->method == PBSIM_METHOD_SAMPLE;  // no model file: set_mut's thresholds only (:5471)
  if (!out || !p || (!model_path && !sample)) {
    *err = "invalid argument";
    return PBSIM_E_INVALID;
  }
  if (p->method != PBSIM_METHOD_QSHMM && p->method != PBSIM_METHOD_ERRHMM && !sample) {
    *err = "method must be qshmm, errhmm or sample";
    return PBSIM_E_INVALID;
  }
  pbsim_host_model *m = new pbsim_host_model();
  std::memset(&m->view, 0, sizeof m->view);
  int rc = sample ? 0 : parse_model(model_path, p->method, m->raw, err);
  if (rc) {
    delete m;
    return rc;
  }
  pbsim_model &v = m->view;
  v.method = p->method;
  v.pass_num = p->pass_num;
  v.len_min = p->len_min;
  v.len_max = p->len_max;
  v.accuracy_mean = p->accuracy_mean;
  std::snprintf(v.id_prefix, sizeof v.id_prefix, "%s", p->id_prefix);
  v.model_acc_min = m->raw.acc_min;
  v.model_acc_max = m->raw.acc_max;
  static thread_local double uni_ep[kAccCells][PBSIM_NQV];
  build_thresholds(v, *p, uni_ep);
  if (sample) {
    v.model_acc_min = v.model_acc_max = 0;
    *out = m;
    return 0;
  }
  rc = build_samplers(*m, *p, err);
  if (rc) {
    delete m;
    return rc;
  }
  build_hmm_rows(*m, uni_ep);
  *out = m;
  return 0;
}

const pbsim_model *pbsim_host_model_get(const pbsim_host_model *m) { return m ? &m->view : nullptr; }

void pbsim_host_model_free(pbsim_host_model *m) { delete m; }

// prob2ssp / ssp_rand_value of simulate_by_*_trans (pbsim.cpp:2504-2528 = :4193-4217): for rank r the start
// fraction 5*j % (j = 0..20) has probability (1/r) / (j+1)^(1 + 1/r), normalised over the 21 outcomes and quantised
// to 1000 table positions like every other table of the reference.
void pbsim_host_ssp_table(int32_t rank_max, uint16_t *ends, uint16_t *mod) {
  for (int j = 0; j < 21; ++j) ends[j] = 0xFFFF;
  mod[0] = 0;
  for (int32_t i = 1; i <= rank_max; ++i) {
    double sum = 0;
    const double value = static_cast<double>(1) / i;
    for (int j = 1; j <= 21; ++j) sum += value / pow(j, (1 + value));
    double total = 0.0;
    long end_wk = 0;
    bool ended = false;
    for (int j = 1; j <= 21; ++j) {
      if (ended) {
        ends[i * 21 + j - 1] = 0xFFFF;
        continue;
      }
      total += (value / pow(j, (1 + value))) / sum;
      end_wk = static_cast<int>(total * 1000 + 0.5);
      if (end_wk > 1000) end_wk = 1000;
      ends[i * 21 + j - 1] = static_cast<uint16_t>(end_wk);
      if (end_wk >= 1000) ended = true;
    }
    mod[i] = static_cast<uint16_t>(end_wk);
  }
}

// the Huffman code and DEFLATE block header the engine's gzip writer uses for a stream with this byte histogram
// (tests: a CPU encoder driven by these tables must produce a stream zlib inflates back to the text)
int pbsim_host_deflate_code(const int64_t hist[256], uint32_t lit[257], uint32_t *hdr_bits, uint32_t *hdr_words,
                            int32_t cap_words) {
  uint64_t h[256];
  for (int i = 0; i < 256; ++i) h[i] = hist[i] > 0 ? static_cast<uint64_t>(hist[i]) : 0;
  pb::GzCode c;
  pb::gz_build_code(h, &c);
  for (int i = 0; i < 256; ++i) lit[i] = c.lit[i];
  lit[256] = c.eob;
  *hdr_bits = c.hdr_bits;
  const int32_t need = static_cast<int32_t>((c.hdr_bits + 31) / 32);
  if (need > cap_words) return PBSIM_E_INVALID;
  for (int32_t i = 0; i < need; ++i) hdr_words[i] = c.hdr[i];
  return 0;
}

// get_sample_inf (pbsim.cpp:1214-1330): the filter in front of --method sample
int pbsim_host_sample_filter(const char *fastq, int64_t bytes, int64_t len_min, int64_t len_max, double accuracy_min,
                             double accuracy_max, char *quals, int64_t *qstart, int64_t qstart_cap, int64_t *n_out,
                             pbsim_sample_stats *st, const char **err) {
  static const char *dummy;
  if (!err) err = &dummy;
  *err = "";
  if (!fastq || bytes < 0 || !quals || !qstart || qstart_cap < 1 || !n_out || !st) {
    *err = "invalid argument";
    return PBSIM_E_INVALID;
  }
  double prob_of[PBSIM_NQV];
  for (int q = 0; q < PBSIM_NQV; ++q) prob_of[q] = std::pow(10, static_cast<double>(q) / -10);  // qc[i].prob (:549)
  std::memset(st, 0, sizeof *st);
  st->len_min = st->len_min_filtered = LONG_MAX;
  std::vector<long> freq_len(static_cast<size_t>(len_max > 0 ? len_max : 0) + 1, 0), freq_acc(100001, 0);
  double accuracy_total = 0;
  int64_t n = 0, out = 0, line_num = 0, pos = 0;
  qstart[0] = 0;
  while (pos < bytes) {
    const char *nl = static_cast<const char *>(std::memchr(fastq + pos, '\n', static_cast<size_t>(bytes - pos)));
    if (!nl) break;  // no line feed: fgets returns the rest, trim() does not count it
    const int64_t len = nl - (fastq + pos);
    const char *line = fastq + pos;
    pos += len + 1;
    if (++line_num < 4) continue;
    line_num = 0;
    if (len > 1000000) {
      *err = "ERROR: fastq is too long. Max acceptable length is 1000000.";
      return PBSIM_E_INVALID;
    }
    st->num++;
    st->len_total += len;
    if (st->num > 100000000) {
      *err = "ERROR: fastq is too many. Max acceptable number is 100000000.";
      return PBSIM_E_INVALID;
    }
    if (len > st->len_max) st->len_max = len;
    if (len < st->len_min) st->len_min = len;
    if (len < len_min || len > len_max) continue;
    double prob = 0.0;
    for (int64_t i = 0; i < len; ++i) {
      const int q = static_cast<unsigned char>(line[i]) - 33;
      if (q < 0 || q >= PBSIM_NQV) {
        *err = "ERROR: quality character outside '!'..'~' in the sample FASTQ.";  // the reference reads qc[] out of bounds
        return PBSIM_E_INVALID;
      }
      prob += prob_of[q];
    }
    const double accuracy = 1.0 - (prob / len);
    if (!(accuracy >= accuracy_min && accuracy <= accuracy_max)) continue;
    if (n + 1 >= qstart_cap) {
      *err = "qstart is too small";
      return PBSIM_E_INVALID;
    }
    accuracy_total += accuracy;
    st->num_filtered++;
    st->len_total_filtered += len;
    freq_len[static_cast<size_t>(len)]++;
    freq_acc[static_cast<size_t>(cxx_int(accuracy * 100000 + 0.5))]++;
    std::memcpy(quals + out, line, static_cast<size_t>(len));
    out += len;
    qstart[++n] = out;
    if (len > st->len_max_filtered) st->len_max_filtered = len;
    if (len < st->len_min_filtered) st->len_min_filtered = len;
  }
  *n_out = n;
  if (st->num_filtered < 1) {
    *err = "ERROR: there is no sample in the valid range of length and accuracy.";
    return PBSIM_E_PARAM;
  }
  st->len_mean_filtered = static_cast<double>(st->len_total_filtered) / st->num_filtered;
  st->accuracy_mean_filtered = accuracy_total / st->num_filtered;
  double variance = 0.0;
  for (int64_t i = 0; i <= len_max; ++i)
    if (freq_len[static_cast<size_t>(i)] > 0) variance += std::pow((st->len_mean_filtered - i), 2) * freq_len[static_cast<size_t>(i)];
  st->len_sd_filtered = std::sqrt(variance / st->num_filtered);
  variance = 0.0;
  for (int64_t i = 0; i <= 100000; ++i)
    if (freq_acc[static_cast<size_t>(i)] > 0) variance += std::pow((st->accuracy_mean_filtered - i * 0.00001), 2) * freq_acc[static_cast<size_t>(i)];
  st->accuracy_sd_filtered = std::sqrt(variance / st->num_filtered);
  return 0;
}

void pbsim_host_hp_del_bias(double opt, const int64_t hpfreq[12], double bias[12]) {
  for (int i = 0; i < 12; ++i) bias[i] = 0.0;
  if (opt == 1) {
    for (int i = 1; i <= 10; ++i) bias[i] = 1;
  } else {
    long sum1 = 0, sum2 = 0;
    for (int i = 1; i <= 10; ++i) {
      bias[i] = 1 + (opt - 1) / 9 * (i - 1);
      sum1 += hpfreq[i] * bias[i];  // long += double: truncates each step, as the reference does (:690)
      sum2 += hpfreq[i];
    }
    const double rate = static_cast<double>(sum2) / sum1;
    for (int i = 1; i <= 10; ++i) bias[i] *= rate;
  }
  // genome.hpfreq[11] (runs >= 11, pbsim.cpp:1046-1058) lies on top of genome.hp_del_bias[0] in the
  // reference build, and hp_del_bias[11] reads the zero padding behind the struct: reproduce both.
  std::memcpy(&bias[0], &hpfreq[11], sizeof(double));
  bias[11] = 0.0;
}

}  // extern "C"
