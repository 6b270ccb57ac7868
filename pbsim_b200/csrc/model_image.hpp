// Device image of a pbsim_model: per-accuracy table blobs in the layout sim_core.cuh reads.
// Built on the host by the engine (engine.cu) when pbsim_cuda_set_model / set_sequence is called.
// Table semantics: see QsView / ErView in sim_core.cuh.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pbsim_cuda.h"

namespace pb {

constexpr uint32_t kQsRows = 51;  // states 0..50, row 0 = init2state

struct AccEntry {          // one per accuracy 0..100, mirrored on the device
  uint32_t blob_off;       // byte offset of the table blob inside ModelImage::blob (16-aligned)
  uint32_t blob_bytes;     // multiple of 16 (TMA bulk copy size)
  uint32_t bias_off;       // errhmm: offset (in uint16 cells) into the per-sequence bias table
  uint32_t nstates;
  uint32_t has_model;      // qshmm: 1 = HMM tables, 0 = freq2qc
  uint32_t init_mod;
  uint32_t freq_mod;
  uint32_t mode;           // errhmm: 0 exact, 1 below model range, 2 above, 3 verbatim (acc 100)
  uint32_t rate_mag;
  uint32_t valid;          // 1 if reads of this accuracy can be simulated
  uint32_t table_acc;      // errhmm: accuracy whose tables drive the chain
  uint32_t seg_ok;         // qshmm: chain couples fast enough for the segment-parallel pass 1
  uint64_t reach;          // qshmm: states reachable from init2state (bit s)
  float rho;               // qshmm: estimated read positions per reference base (segment provisioning)
  uint32_t pad2;
};

// blob layouts (all sections 16-byte aligned)
struct QsBlobLayout {
  static constexpr uint32_t t2_off = 0;                                   // uint32[51*100]
  static constexpr uint32_t emis_off = ((kQsRows * 100 * 4 + 15) / 16) * 16;  // uint8[51*100]
  static constexpr uint32_t tmod_off = emis_off + ((kQsRows * 100 + 15) / 16) * 16;  // uint8[51] (+pad to 64)
  static constexpr uint32_t emodv_off = tmod_off + 64;                               // uint8[51] (+pad to 64)
  static constexpr uint32_t bytes = emodv_off + 64;
  static constexpr uint32_t freq_bytes = 1008;                            // uint8[1000] padded
};

#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t er_blob_bytes(uint32_t nst, uint32_t *t2_off, uint32_t *emis_off, uint32_t *emod_off) {
  const uint32_t rows = nst + 1;
  uint32_t o = 0;
  *t2_off = o;   o += ((rows * 1000 * 2 + 15) / 16) * 16;
  *emis_off = o; o += ((rows * 1000 + 15) / 16) * 16;
  *emod_off = o; o += ((rows * 4 + 15) / 16) * 16;  // emod[rows] then tmod[rows] (uint16 each)
  return o;
}

// Does the chain of this accuracy forget its past quickly?  32 trials of the grand coupling (all reachable
// states driven by the same uniform draws): segment-parallel pass 1 is enabled when every trial coalesces
// within 2048 steps and the median is below 512.  Only a performance switch: results are identical either way
// (the backward search itself never fails, it just gets longer).
inline uint32_t coupling_screen(const pbsim_hmm_row &r, uint64_t reach) {
  uint64_t x = 0x9E3779B97F4A7C15ull;
  int below = 0;
  int times[32];
  const int res = r.resolution;
  for (int trial = 0; trial < 32; ++trial) {
    uint64_t mask = reach;
    int t = 0;
    for (; t < 2048 && (mask & (mask - 1)); ++t) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;  // xorshift64
      const uint32_t u = (uint32_t)(x >> 32);
      uint64_t next = 0;
      for (uint64_t m = mask; m; m &= m - 1) {
        const int s = __builtin_ctzll(m);
        if (s < 1 || s > r.nstates) { next |= 1ull << s; continue; }
        const int tm = r.tran_mod[s] < 1 ? 1 : r.tran_mod[s];
        next |= 1ull << r.tran[s * res + (int)(((uint64_t)u * (uint32_t)tm) >> 32)];
      }
      mask = next;
    }
    if (mask & (mask - 1)) return 0;
    if (t < 512) ++below;
    times[trial] = t;
  }
  if (below < 16) return 0;
  // first backward window: a bit above the ~90th percentile of the observed coalescence times, power of two
  for (int i = 0; i < 32; ++i)
    for (int j = i + 1; j < 32; ++j)
      if (times[j] < times[i]) { const int tt = times[i]; times[i] = times[j]; times[j] = tt; }
  uint32_t w = 16;
  while (w < (uint32_t)times[28] + 8u && w < 2048u) w *= 2u;
  return w;
}

// states reachable from the init row (bit s)
inline uint64_t reachable_states(const pbsim_hmm_row &r) {
  uint64_t reach = 0, frontier = 0;
  const int res = r.resolution;
  for (int k = 0; k < r.init_mod && k < res; ++k) frontier |= 1ull << r.init[k];
  while (frontier) {
    const int s2 = __builtin_ctzll(frontier);
    frontier &= frontier - 1;
    if (reach >> s2 & 1ull) continue;
    reach |= 1ull << s2;
    if (s2 >= 1 && s2 <= r.nstates)
      for (int k = 0; k < r.tran_mod[s2] && k < res; ++k) {
        const int nx = r.tran[s2 * res + k];
        if (!(reach >> nx & 1ull)) frontier |= 1ull << nx;
      }
  }
  return reach;
}

// Estimated read positions per consumed reference base for one accuracy (Monte Carlo over the quantised
// tables, 1<<15 positions): only used to provision segments, never affects results.
inline float estimate_rho(const pbsim_model &m, const pbsim_hmm_row &r) {
  uint64_t x = 0xD1B54A32D192ED03ull;
  auto u32 = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (uint32_t)(x >> 32); };
  auto pick = [&](uint32_t mod) { return (uint32_t)(((uint64_t)u32() * (mod < 1 ? 1 : mod)) >> 32); };
  uint64_t ref = 0;
  const int N = 1 << 15;
  int state = 0;
  for (int p = 0; p < N; ++p) {
    int qv;
    if (r.exists) {
      state = (p == 0 || state < 1 || state > r.nstates) ? r.init[pick(r.init_mod)] : r.tran[state * 100 + pick(r.tran_mod[state])];
      if (state < 1 || state > r.nstates) state = r.init[pick(r.init_mod)];
      qv = r.emis[state * 100 + pick(r.emis_mod[state])];
    } else {
      qv = r.freq[pick(r.freq_mod)];
    }
    const uint32_t e = pick(1000000);
    if (!(e >= (uint32_t)m.sub_thre[qv] && e < (uint32_t)m.ins_thre[qv])) ++ref;  // not an insertion
    int guard = 0;
    while (pick(1000000) < (uint32_t)m.del_thre[qv] && ++guard < 64) ++ref;
  }
  return ref ? (float)((double)N / (double)ref) : 1.0f;
}

// errhmm: estimated alignment columns per consumed reference base (segment provisioning only)
inline float estimate_rho_err(const pbsim_hmm_row &r, uint32_t mode, uint32_t rate_mag) {
  uint64_t x = 0xA0761D6478BD642Full;
  auto u32 = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (uint32_t)(x >> 32); };
  auto pick = [&](uint32_t mod) { return (uint32_t)(((uint64_t)u32() * (mod < 1 ? 1 : mod)) >> 32); };
  const int N = 1 << 15;
  uint64_t ref = 0;
  int state = 0;
  for (int c = 0; c < N; ++c) {
    state = (c == 0 || state < 1 || state > r.nstates) ? r.init[pick(r.init_mod)] : r.tran[state * 1000 + pick(r.tran_mod[state])];
    if (state < 1 || state > r.nstates) state = r.init[pick(r.init_mod)];
    uint32_t kind;
    if (pick(1000) + 1 <= (uint32_t)r.emis_del[state]) kind = 3;
    else kind = r.emis_mod[state] == 0 ? pick(3) : r.emis[state * 1000 + pick(r.emis_mod[state])];
    const uint32_t mag = pick(100) + 1;
    if (mode == 1 && kind == 0 && mag <= rate_mag) kind = pick(3) + 1;
    if (mode == 2 && kind != 0 && mag <= rate_mag) kind = 0;
    if (kind != 2) ++ref;
  }
  return ref ? (float)((double)N / (double)ref) : 1.0f;
}

struct ModelImage {
  int method = 0;
  std::vector<uint8_t> blob;
  AccEntry acc[PBSIM_NACC];
  // per-sequence, bias dependent
  std::vector<uint32_t> qs_thr;      // [94*4]
  std::vector<uint32_t> qs_thr_hp;   // [94*12]
  // PHILOX mode: the same thresholds on the 32-bit scale T32(t) (sim_core.cuh), and the per-quality record of
  // the position-parallel kernels {T32(sub), T32(ins), T32(del), error probability in 2^-26 fixed point}
  std::vector<uint32_t> qs_thr32, qs_thr_hp32, qs_fast;
  std::vector<uint16_t> er_bias;     // per table accuracy: edel[rows] then edel_hp[rows*12]
  std::string error;
  bool uniform_bias = true;          // hp_del_bias[1..10] all exactly 1

  bool build(const pbsim_model &m) {
    method = m.method;
    blob.clear();
    std::memset(acc, 0, sizeof acc);
    uint32_t bias_cells = 0;
    for (int a = 0; a < PBSIM_NACC; ++a) acc[a].table_acc = a;
    if (method == PBSIM_METHOD_SAMPLE) return true;  // no HMM: the thresholds of apply_bias are all there is
    if (method == PBSIM_METHOD_QSHMM) {
      for (int a = m.acc_lo; a <= m.acc_hi; ++a) {
        if (a < 0 || a >= PBSIM_NACC) continue;
        const pbsim_hmm_row &r = m.rows[a];
        AccEntry &e = acc[a];
        e.blob_off = (uint32_t)blob.size();
        if (r.exists) {
          if (!r.tran || !r.emis || !r.init || r.nstates + 1 > (int)kQsRows || r.resolution != 100) {
            error = "qshmm row malformed";
            return false;
          }
          blob.resize(blob.size() + QsBlobLayout::bytes, 0);
          uint8_t *b = blob.data() + e.blob_off;
          uint32_t *t2 = reinterpret_cast<uint32_t *>(b + QsBlobLayout::t2_off);
          uint8_t *em = b + QsBlobLayout::emis_off;
          // an entry carries the ROW OFFSET (state*100) and both moduli of the state it leads to:
          // one shared-memory load per chain step, no multiply for the next address
          auto entry = [&](int s) -> uint32_t {
            int tm = (s >= 1 && s <= r.nstates) ? r.tran_mod[s] : 1;
            int emd = (s >= 1 && s <= r.nstates) ? r.emis_mod[s] : 1;
            tm = tm < 1 ? 1 : (tm > 255 ? 255 : tm);
            emd = emd < 1 ? 1 : (emd > 255 ? 255 : emd);
            const int st = (s >= 0 && s <= r.nstates) ? s : 0;
            return (uint32_t)(st * 100) | ((uint32_t)tm << 16) | ((uint32_t)emd << 24);
          };
          for (int k = 0; k < r.init_mod && k < 100; ++k) t2[k] = entry(r.init[k]);
          for (int s = 1; s <= r.nstates; ++s) {
            for (int k = 0; k < 100; ++k) {
              t2[s * 100 + k] = entry(r.tran[s * 100 + k]);
              em[s * 100 + k] = r.emis[s * 100 + k];
            }
          }
          uint8_t *tmodv = b + QsBlobLayout::tmod_off, *emodv = b + QsBlobLayout::emodv_off;
          for (int s2 = 0; s2 <= 50; ++s2) {
            const uint32_t en = entry(s2);
            tmodv[s2] = (uint8_t)((en >> 16) & 0xFFu);
            emodv[s2] = (uint8_t)(en >> 24);
          }
          // reachable closure from the init row, then a Monte-Carlo screen of the grand coupling time
          const uint64_t reach = reachable_states(r);
          e.reach = reach;
          e.seg_ok = coupling_screen(r, reach);  // 0, or the first backward-coupling window
          e.rho = estimate_rho(m, r);
          e.blob_bytes = QsBlobLayout::bytes;
          e.has_model = 1;
          e.nstates = (uint32_t)r.nstates;
          e.init_mod = (uint32_t)(r.init_mod < 1 ? 1 : r.init_mod);
          e.valid = 1;
        } else {
          if (!r.freq || r.freq_mod < 1) {
            error = "qshmm freq row missing";
            return false;
          }
          blob.resize(blob.size() + QsBlobLayout::freq_bytes, 0);
          std::memcpy(blob.data() + e.blob_off, r.freq, (size_t)r.freq_mod);
          e.blob_bytes = QsBlobLayout::freq_bytes;
          e.seg_ok = 1;  // no chain at all: positions are independent
          e.rho = estimate_rho(m, r);
          e.has_model = 0;
          e.freq_mod = (uint32_t)r.freq_mod;
          e.valid = 1;
        }
      }
    } else {
      // blobs for modelled accuracies inside the sampler's range
      for (int a = m.acc_lo; a <= m.acc_hi; ++a) {
        if (a < 0 || a >= PBSIM_NACC) continue;
        const pbsim_hmm_row &r = m.rows[a];
        if (!r.exists || r.nstates < 1 || !r.tran) continue;
        if (r.resolution != 1000 || r.nstates > 50) {
          error = "errhmm row malformed";
          return false;
        }
        AccEntry &e = acc[a];
        uint32_t t2o, emo, emodo;
        const uint32_t bytes = er_blob_bytes((uint32_t)r.nstates, &t2o, &emo, &emodo);
        e.blob_off = (uint32_t)blob.size();
        blob.resize(blob.size() + bytes, 0);
        uint8_t *b = blob.data() + e.blob_off;
        uint16_t *t2 = reinterpret_cast<uint16_t *>(b + t2o);
        uint8_t *em = b + emo;
        uint16_t *emod = reinterpret_cast<uint16_t *>(b + emodo);
        auto tmod = [&](int s) -> uint32_t {
          int v = (s >= 1 && s <= r.nstates) ? r.tran_mod[s] : 1;
          return (uint32_t)(v < 1 ? 1 : (v > 1000 ? 1000 : v));
        };
        for (int k = 0; k < r.init_mod && k < 1000; ++k) t2[k] = (uint16_t)(r.init[k] | (tmod(r.init[k]) << 6));
        for (int s = 1; s <= r.nstates; ++s) {
          for (int k = 0; k < 1000; ++k) {
            const uint8_t nx = r.tran[s * 1000 + k];
            t2[s * 1000 + k] = (uint16_t)(nx | (tmod(nx) << 6));
            em[s * 1000 + k] = r.emis[s * 1000 + k];
          }
          emod[s] = (uint16_t)(r.emis_mod[s] < 0 ? 0 : r.emis_mod[s]);
          emod[(r.nstates + 1) + s] = (uint16_t)tmod(s);  // per-state transition modulus (chain-only prepass)
        }
        e.rho = 1.0f;
        e.blob_bytes = bytes;
        e.nstates = (uint32_t)r.nstates;
        e.init_mod = (uint32_t)(r.init_mod < 1 ? 1 : r.init_mod);
        e.bias_off = bias_cells;
        bias_cells += (((uint32_t)(r.nstates + 1) * 13u + 7u) / 8u) * 8u;  // keep regions 16-byte aligned
        e.valid = 1;
      }
      // every accuracy the sampler can return borrows the tables of a modelled one (:3852-3926)
      for (int a = m.acc_lo; a <= m.acc_hi; ++a) {
        if (a < 0 || a >= PBSIM_NACC) continue;
        AccEntry &e = acc[a];
        if (a == 100) {  // verbatim copy (:3837-3845)
          e.valid = 1;
          e.mode = 3;
          e.blob_bytes = 0;
          continue;
        }
        if (m.rows[a].exists) {
          e.mode = 0;
          continue;
        }
        int ta;
        if (a < m.model_acc_min) {
          ta = m.model_acc_min;
          e.mode = 1;
          e.rate_mag = (uint32_t)(int)((double)(m.model_acc_min - a) / m.model_acc_min * 100);
        } else {
          ta = m.model_acc_max;
          e.mode = 2;
          e.rate_mag = (uint32_t)(int)((double)(a - m.model_acc_max) / (100 - m.model_acc_max) * 100);
        }
        if (ta < 0 || ta >= PBSIM_NACC || !acc[ta].valid || acc[ta].mode == 3) {
          // the reference would read tables it never built (undefined); refuse instead
          e.valid = 0;
          continue;
        }
        const uint32_t mode = e.mode, mag = e.rate_mag;
        e = acc[ta];
        e.mode = mode;
        e.rate_mag = mag;
        e.table_acc = (uint32_t)ta;
      }
      for (int a = m.acc_lo; a <= m.acc_hi; ++a) {
        if (a < 0 || a >= PBSIM_NACC) continue;
        AccEntry &e = acc[a];
        if (!e.valid || e.mode == 3) continue;
        e.rho = estimate_rho_err(m.rows[e.table_acc], e.mode, e.rate_mag);
        e.reach = reachable_states(m.rows[e.table_acc]);
        e.seg_ok = coupling_screen(m.rows[e.table_acc], e.reach);
      }
      er_bias.assign(bias_cells, 0);
    }
    return true;
  }

  // thresholds that depend on genome.hp_del_bias (per sequence)
  void apply_bias(const pbsim_model &m, const double bias[12]) {
    uniform_bias = true;
    for (int h = 1; h <= 10; ++h)
      if (bias[h] != 1.0) uniform_bias = false;
    if (method == PBSIM_METHOD_QSHMM || method == PBSIM_METHOD_SAMPLE) {
      qs_thr.assign(PBSIM_NQV * 4, 0);
      qs_thr_hp.assign(PBSIM_NQV * 12, 0);
      for (int q = 0; q < PBSIM_NQV; ++q) {
        uint32_t mx = 0;
        for (int h = 0; h < 12; ++h) {
          // reference test: rand_value < del_thre[qv] * hp_del_bias[hp]   (long < double, :2272)
          const double x = (double)m.del_thre[q] * bias[h];
          double c = std::ceil(x);
          if (!(c >= 0)) c = 0;
          if (c > 1000000.0) c = 1000000.0;
          const uint32_t t = (uint32_t)c;
          qs_thr_hp[q * 12 + h] = t;
          if (t > mx) mx = t;
        }
        qs_thr[q * 4 + 0] = (uint32_t)m.sub_thre[q];
        qs_thr[q * 4 + 1] = (uint32_t)m.ins_thre[q];
        qs_thr[q * 4 + 2] = mx;
        qs_thr[q * 4 + 3] = qs_thr_hp[q * 12 + 0];
      }
      auto t32 = [](uint32_t t) -> uint32_t {
        if (t >= 1000000u) return 0xFFFFFFFFu;
        return (uint32_t)(((uint64_t)t * 4294967296ull + 999999ull) / 1000000ull);
      };
      qs_thr32.resize(qs_thr.size());
      qs_thr_hp32.resize(qs_thr_hp.size());
      qs_fast.assign(PBSIM_NQV * 4, 0);
      for (size_t i = 0; i < qs_thr.size(); ++i) qs_thr32[i] = t32(qs_thr[i]);
      for (size_t i = 0; i < qs_thr_hp.size(); ++i) qs_thr_hp32[i] = t32(qs_thr_hp[i]);
      for (int q = 0; q < PBSIM_NQV; ++q) {
        qs_fast[q * 4 + 0] = qs_thr32[q * 4 + 0];
        qs_fast[q * 4 + 1] = qs_thr32[q * 4 + 1];
        qs_fast[q * 4 + 2] = qs_thr32[q * 4 + 2];
        qs_fast[q * 4 + 3] = (uint32_t)std::llround(m.qc_prob[q] * 67108864.0);  // 2^26
      }
    } else {
      for (int a = 0; a < PBSIM_NACC; ++a) {
        const AccEntry &e = acc[a];
        if (!e.valid || e.mode == 3 || e.table_acc != (uint32_t)a) continue;
        const pbsim_hmm_row &r = m.rows[a];
        uint16_t *edel = er_bias.data() + e.bias_off;
        uint16_t *edel_hp = edel + (e.nstates + 1);
        for (uint32_t s = 0; s <= e.nstates; ++s) {
          uint32_t mx = 0;
          for (int h = 0; h < 12; ++h) {
            // reference test: index <= emis2del[state] * hp_del_bias[hp]   (long <= double, :3862)
            const double x = (s >= 1 ? (double)r.emis_del[s] : 0.0) * bias[h];
            double f = std::floor(x);
            if (!(f >= 0)) f = 0;
            if (f > 65535.0) f = 65535.0;
            const uint32_t t = (uint32_t)f;
            edel_hp[s * 12 + h] = (uint16_t)t;
            if (t > mx) mx = t;
          }
          edel[s] = (uint16_t)mx;
        }
      }
    }
  }
};

}  // namespace pb
