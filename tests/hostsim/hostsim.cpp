// TEST HARNESS (not product): compiles the engine's per-read core (pbsim_b200/csrc/sim_core.cuh)
// and table image builder (model_image.hpp) as plain C++ and runs them sequentially, read by
// read, so that the pass-1 logic (planning, chains, event encoding, both draw sources) can be
// checked against the oracle on a box without a GPU.  The product never runs this code path:
// libpbsim_cuda executes the same header only inside CUDA kernels.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/pbsim_cuda.h"
#include "../../pbsim_b200/csrc/model_image.hpp"
#include "../../pbsim_b200/csrc/sample_plan.hpp"
#include "../../pbsim_b200/csrc/sim_core.cuh"

struct HostSimOut {
  std::vector<int64_t> info;     // 12 per subread: read_id, pass, acc, offset, wlen, rlen, ncol, minus, n_entries, ev_off, draw_start, overflow
  std::vector<uint32_t> counts;  // 3 per subread: nsub, nins, ndel
  std::vector<double> accuracy;
  std::vector<uint8_t> events;   // qshmm: uint16 entries (little endian), errhmm: uint8 entries
  std::vector<pb::Ckpt> ckpts;
  std::vector<int64_t> ck_off;
};

static HostSimOut g_out;
static int use_fast = 1, use_seg = 0, seg_min_len = 2048;
static long seg_reads = 0, seg_fallbacks = 0;
extern "C" void hostsim_use_fast(int v) { use_fast = v; }
extern "C" void hostsim_use_segments(int v, int min_len) { use_seg = v; seg_min_len = min_len; }
// > 0: the states in front of the segments come from chain chunks of this many segments, as k_chain_chunk computes
// them (coupling at the chunk's first position, then the chain-only walk); 0: coupling per segment / whole-read walk
static int chain_chunk = 0;
extern "C" void hostsim_set_chain_chunk(int g) { chain_chunk = g; }
extern "C" long hostsim_seg_reads() { return seg_reads; }
extern "C" long hostsim_seg_fallbacks() { return seg_fallbacks; }

// the segment-parallel pass 1 as the kernels run it (k_sim_seg + k_find_end), sequentially on the host;
// returns false if the read must fall back to the sequential path
static bool run_segmented(const pb::QsView &T, const pb::QsSegAux &A, uint32_t seed, uint32_t seq_num, uint32_t read_id,
                          uint32_t pass, uint32_t wlen, float rho, bool seg_ok, uint32_t first_window, const pb::HpProbe &hp,
                          std::vector<uint8_t> &events,
                          size_t ev_off, std::vector<pb::Ckpt> &ckpts, size_t ck_base, pb::SubreadResult &res) {
  pb::PhiloxKeys K;
  K.init(seed, seq_num);
  const uint32_t n_seg = pb::qshmm_segments_for(wlen, rho);
  std::vector<uint16_t> slots((size_t)n_seg * PB_SEG_STRIDE + 16, 0);
  std::vector<pb::SegResult> seg(n_seg);
  if (T.has_model) {  // k_chain_chunk: the quality pass, one "thread" per chunk of segments
    const uint32_t cc = chain_chunk > 0 ? (uint32_t)chain_chunk : 1u;
    const uint32_t n_ch = seg_ok ? (n_seg + cc - 1u) / cc : 1u;
    const uint32_t per = n_ch == 1u ? n_seg : cc;
    for (uint32_t c = 0; c < n_ch; ++c) {
      const uint32_t k_from = c * per, k_to = std::min(k_from + per, n_seg);
      uint32_t row = 0, mod = T.init_mod, emod = 1;
      if (k_from > 0) pb::qshmm_segment_start(T, A, K, read_id, pass, k_from * PB_TILE, first_window ? first_window : 512u, row, mod, emod);
      pb::qshmm_quality_range(T, K, read_id, pass, row, mod, emod, k_from, k_to, slots.data());
    }
  }
  for (uint32_t k = 0; k < n_seg; ++k)  // k_sim_seg: the error pass, one "warp" per segment
    pb::qshmm_error_segment(T, K, read_id, pass, k * PB_TILE, k == 0, slots.data() + (size_t)k * PB_SEG_STRIDE, seg[k]);
  std::vector<pb::Ckpt> ck(n_seg);
  pb::SegRead sr;
  pb::qshmm_finish_segmented(slots.data(), seg.data(), n_seg, wlen, T.fast, hp, ck.data(), sr);
  if (sr.flags) return false;
  if (hp.enabled) {
    // what pass 2 does in PHILOX mode: the 4-way choice of a substitution on a non-ACGT base is recomputed from
    // the position's own Philox block (the segment could not know the base)
    for (uint32_t k = 0; k < sr.n_tiles; ++k) {
      uint32_t R = ck[k].ref, P = ck[k].read;
      uint16_t *e = slots.data() + (size_t)k * PB_SEG_STRIDE;
      for (uint32_t i = 0; i < ck[k].pad; ++i) {
        const uint32_t v = e[i], kind = (v >> 7) & 3u;
        if (kind == 3u) { R += (v & 0x7Fu) | ((v >> 9) << 7); continue; }
        if (kind == PB_KIND_SUB && hp.win.nonacgt(R)) {
          uint32_t x, y;
          pb::error_words_at(K, read_id, pass << 16, P, x, y);
          e[i] = (uint16_t)((v & ~(7u << 9)) | (((y >> 3) & 3u) << 9));
        }
        R += (kind == PB_KIND_INS ? 0u : 1u) + (v >> 12);
        ++P;
      }
    }
  }
  // linearise the tiles for the Python expander
  size_t total = 0;
  for (uint32_t k = 0; k < sr.n_tiles; ++k) total += ck[k].pad;
  events.resize(ev_off + total * 2);
  uint16_t *dst = reinterpret_cast<uint16_t *>(events.data() + ev_off);
  for (uint32_t k = 0; k < sr.n_tiles; ++k) {
    memcpy(dst, slots.data() + (size_t)k * PB_SEG_STRIDE, (size_t)ck[k].pad * 2);
    dst += ck[k].pad;
  }
  ckpts.resize(ck_base + sr.n_tiles);
  for (uint32_t k = 0; k < sr.n_tiles; ++k) ckpts[ck_base + k] = ck[k];
  res.n_entries = (uint32_t)total; res.rlen = sr.rlen; res.ncol = sr.ncol; res.nsub = sr.nsub; res.nins = sr.nins;
  res.ndel = sr.ndel; res.overflow = 0; res.accuracy = sr.accuracy;
  return true;
}

// --method sample, the speculative pass of a long read as the kernels run it: the qualities of the segments' positions
// come from the pool entry (k_sim_seg<true>), the error pass is qshmm's, and the read ends where its window is used up
// or where it is as long as its quality string (k_find_end with sample != 0).  Returns false: fall back.
static bool run_segmented_sample(pb::QsView T, uint32_t seed, uint32_t seq_num, uint32_t read_id, uint32_t len,
                                 const uint8_t *quals, const pb::HpProbe &hp, std::vector<uint8_t> &events, size_t ev_off,
                                 std::vector<pb::Ckpt> &ckpts, size_t ck_base, pb::SubreadResult &res) {
  pb::PhiloxKeys K;
  K.init(seed, seq_num);
  const uint32_t n_seg = (len + PB_TILE - 1u) / PB_TILE;
  std::vector<uint16_t> slots((size_t)n_seg * PB_SEG_STRIDE + 16, 0);
  std::vector<pb::SegResult> seg(n_seg);
  T.has_model = 1;  // "the qualities are in the slot"
  for (uint32_t k = 0; k < n_seg; ++k) {
    uint16_t *slot = slots.data() + (size_t)k * PB_SEG_STRIDE;
    for (uint32_t j = 0; j < PB_TILE; ++j) {
      const uint32_t p = k * PB_TILE + j;
      slot[j] = p < len ? (uint16_t)(quals[p] - 33u) : 0u;
    }
    pb::qshmm_error_segment(T, K, read_id, 0u, k * PB_TILE, k == 0, slot, seg[k]);
  }
  std::vector<pb::Ckpt> ck(n_seg);
  pb::SegRead sr;
  pb::qshmm_finish_segmented(slots.data(), seg.data(), n_seg, len, T.fast, hp, ck.data(), sr, len);
  if (sr.flags) return false;
  if (hp.enabled) {  // pass 2's re-derivation of the 4-way choice of a substitution on a non-ACGT base
    for (uint32_t k = 0; k < sr.n_tiles; ++k) {
      uint32_t R = ck[k].ref, P = ck[k].read;
      uint16_t *e = slots.data() + (size_t)k * PB_SEG_STRIDE;
      for (uint32_t i = 0; i < ck[k].pad; ++i) {
        const uint32_t v = e[i], kind = (v >> 7) & 3u;
        if (kind == 3u) { R += (v & 0x7Fu) | ((v >> 9) << 7); continue; }
        if (kind == PB_KIND_SUB && hp.win.nonacgt(R)) {
          uint32_t x, y;
          pb::error_words_at(K, read_id, 0u, P, x, y);
          e[i] = (uint16_t)((v & ~(7u << 9)) | (((y >> 3) & 3u) << 9));
        }
        R += (kind == PB_KIND_INS ? 0u : 1u) + (v >> 12);
        ++P;
      }
    }
  }
  size_t total = 0;
  for (uint32_t k = 0; k < sr.n_tiles; ++k) total += ck[k].pad;
  events.resize(ev_off + total * 2);
  uint16_t *dst = reinterpret_cast<uint16_t *>(events.data() + ev_off);
  for (uint32_t k = 0; k < sr.n_tiles; ++k) {
    memcpy(dst, slots.data() + (size_t)k * PB_SEG_STRIDE, (size_t)ck[k].pad * 2);
    dst += ck[k].pad;
  }
  ckpts.resize(ck_base + sr.n_tiles);
  for (uint32_t k = 0; k < sr.n_tiles; ++k) ckpts[ck_base + k] = ck[k];
  res.n_entries = (uint32_t)total; res.rlen = sr.rlen; res.ncol = sr.ncol; res.nsub = sr.nsub; res.nins = sr.nins;
  res.ndel = sr.ndel; res.overflow = 0; res.accuracy = sr.accuracy;
  return true;
}

extern "C" {

// returns number of subreads, <0 on error
long hostsim_run(const pbsim_model *m, const uint8_t *ascii_upper, const int16_t *hp, long glen, int seq_num,
                 const double bias[12], int rng_mode, uint32_t seed, const int32_t *draws, long ndraws,
                 long long len_quota, long max_reads) {
  pb::ModelImage img;
  if (!img.build(*m)) { fprintf(stderr, "hostsim: %s\n", img.error.c_str()); return -1; }
  img.apply_bias(*m, bias);
  std::vector<uint8_t> hp4((size_t)glen / 2 + 2, 0);
  for (long i = 0; i < glen; ++i) hp4[i >> 1] |= (uint8_t)((hp[i] & 15) << ((i & 1) * 4));
  bool any_exc = !img.uniform_bias;
  std::vector<uint8_t> exc((size_t)glen, 0);
  for (long i = 0; i < glen; ++i) {
    const uint8_t c = ascii_upper[i];
    const bool acgt = (c == 'A' || c == 'C' || c == 'G' || c == 'T');
    exc[i] = (!acgt) || (bias[hp[i] & 15] != 1.0);
  }
  std::vector<uint32_t> xm(((size_t)glen >> 15) + 2, 0);  // 1 bit per 1024-base block with an exceptional base
  for (long i = 0; i < glen; ++i)
    if (exc[i]) xm[(i >> 10) >> 5] |= 1u << ((i >> 10) & 31);
  uint8_t bias_one[12];
  for (int h = 0; h < 12; ++h) bias_one[h] = (bias[h] == 1.0) ? 1 : 0;
  bias_one[0] = 1;
  g_out = HostSimOut();
  pb::PlanTables pt;
  pt.prob2len = m->prob2len;
  pt.prob2acc = m->prob2accuracy;
  pt.len_rand_value = (uint32_t)m->len_rand_value;
  pt.acc_rand_value = (uint32_t)m->accuracy_rand_value;
  pt.len_min = (uint32_t)m->len_min;

  long long len_total = 0;
  long read_id = 0;
  int64_t cursor = 0;
  while (len_total < len_quota && (max_reads <= 0 || read_id < max_reads)) {
    ++read_id;
    pb::PhiloxDraw pd;
    pb::ReplayDraw rd;
    pb::ReadPlan plan;
    const int64_t read_start = cursor;
    if (rng_mode == PBSIM_RNG_PHILOX) {
      pd.ph.k0 = seed; pd.ph.k1 = (uint32_t)seq_num; pd.read_id = (uint32_t)read_id; pd.pass = 0;
      plan = pb::plan_read(pt, pd, (uint32_t)glen, len_quota - len_total);
    } else {
      rd.log = draws; rd.cur = cursor; rd.end = ndraws; rd.start = cursor;
      plan = pb::plan_read(pt, rd, (uint32_t)glen, len_quota - len_total);
      cursor = rd.cur;
    }
    const pb::AccEntry &ae = img.acc[plan.acc];
    if (!ae.valid) { fprintf(stderr, "hostsim: accuracy %u has no tables\n", plan.acc); return -2; }
    const uint32_t minus = (read_id % 2 == 0);
    pb::WindowRef win;
    win.ascii = ascii_upper; win.hp4 = hp4.data(); win.offset = plan.offset; win.wlen = plan.wlen; win.minus = minus;
    bool slow = any_exc;
    for (uint32_t i = 0; i < plan.wlen && !slow; ++i) slow = exc[plan.offset + i];
    for (int pass = 0; pass < m->pass_num; ++pass) {
      const uint32_t cap = plan.wlen * 4 + 4096;
      pb::SubreadResult res;
      const size_t ck_base = g_out.ckpts.size();
      g_out.ckpts.resize(ck_base + cap / PB_TILE + 2);
      const int64_t draw_start = (pass == 0) ? read_start : cursor;
      size_t ev_off = g_out.events.size();
      ev_off = (ev_off + 15) / 16 * 16;
      if (m->method == PBSIM_METHOD_QSHMM) {
        g_out.events.resize(ev_off + (size_t)cap * 2);
        pb::QsView T;
        const uint8_t *b = img.blob.data() + ae.blob_off;
        T.t2 = reinterpret_cast<const uint32_t *>(b + pb::QsBlobLayout::t2_off);
        T.emis = b + pb::QsBlobLayout::emis_off;
        T.freq = b;
        T.has_model = ae.has_model; T.init_mod = ae.init_mod; T.freq_mod = ae.freq_mod;
        T.thr = reinterpret_cast<const pb::QsThr *>(img.qs_thr.data()); T.thr_hp = img.qs_thr_hp.data(); T.qc_prob = m->qc_prob;
        T.thr32 = reinterpret_cast<const pb::QsThr *>(img.qs_thr32.data()); T.thr_hp32 = img.qs_thr_hp32.data();
        T.fast = reinterpret_cast<const pb::QsFast *>(img.qs_fast.data());
        pb::PhiloxDrawQ pq;
        pq.ph.k0 = seed; pq.ph.k1 = (uint32_t)seq_num; pq.read_id = (uint32_t)read_id; pq.pass = (uint32_t)pass;
        pb::QsSink sink;
        sink.init(reinterpret_cast<uint16_t *>(g_out.events.data() + ev_off), g_out.ckpts.data() + ck_base, cap);
        bool seg_done = false;
        if (rng_mode == PBSIM_RNG_PHILOX && (!slow || img.uniform_bias) && use_seg && (int)plan.wlen >= seg_min_len) {
          pb::HpProbe hpp;
          hpp.enabled = slow ? 1u : 0u;
          hpp.win = win;
          hpp.xm = xm.data();
          hpp.bias_one = bias_one;
          pb::QsSegAux A;
          A.tmod = b + pb::QsBlobLayout::tmod_off;
          A.emodv = b + pb::QsBlobLayout::emodv_off;
          A.reach = ae.reach;
          ++seg_reads;
          seg_done = run_segmented(T, A, seed, (uint32_t)seq_num, (uint32_t)read_id, (uint32_t)pass, plan.wlen, ae.rho, ae.seg_ok != 0, ae.seg_ok, hpp,
                                   g_out.events, ev_off,
                                   g_out.ckpts, ck_base, res);
          if (!seg_done) ++seg_fallbacks;
        }
        if (seg_done) {
        } else if (rng_mode == PBSIM_RNG_PHILOX && !slow && use_fast) {
          pb::PhiloxKeys K;
          K.init(seed, (uint32_t)seq_num);
          pb::qshmm_simulate_fast(T, K, (uint32_t)read_id, (uint32_t)pass, plan.wlen,
                                  reinterpret_cast<uint16_t *>(g_out.events.data() + ev_off), g_out.ckpts.data() + ck_base, cap, res);
        } else if (rng_mode == PBSIM_RNG_PHILOX) { pb::qshmm_simulate(T, pq, win, slow, plan.wlen, sink, res); }
        else { rd.cur = cursor; pb::qshmm_simulate(T, rd, win, slow, plan.wlen, sink, res); cursor = rd.cur; }
        if (!seg_done) g_out.events.resize(ev_off + (size_t)res.n_entries * 2);
      } else {
        g_out.events.resize(ev_off + (size_t)cap);
        pb::ErView T;
        std::memset(&T, 0, sizeof T);
        T.mode = ae.mode; T.rate_mag = ae.rate_mag; T.init_mod = ae.init_mod;
        if (ae.mode != 3) {
          uint32_t t2o, emo, emodo;
          pb::er_blob_bytes(ae.nstates, &t2o, &emo, &emodo);
          const uint8_t *b = img.blob.data() + ae.blob_off;
          T.t2 = reinterpret_cast<const uint16_t *>(b + t2o);
          T.emis = b + emo;
          T.emod = reinterpret_cast<const uint16_t *>(b + emodo);
          T.edel = img.er_bias.data() + ae.bias_off;
          T.edel_hp = T.edel + (ae.nstates + 1);
        }
        pb::ErSink sink;
        sink.init(g_out.events.data() + ev_off, g_out.ckpts.data() + ck_base, cap);
        bool seg_done = false;
        if (rng_mode == PBSIM_RNG_PHILOX && (!slow || img.uniform_bias) && use_seg && ae.mode != 3 &&
            (int)plan.wlen >= seg_min_len) {
          // segment-parallel errhmm as the kernels run it: chain-only prepass, segments, find_end
          pb::HpProbe hpp;
          hpp.enabled = slow ? 1u : 0u;
          hpp.win = win; hpp.xm = xm.data(); hpp.bias_one = bias_one;
          pb::PhiloxKeys K;
          K.init(seed, (uint32_t)seq_num);
          const uint32_t n_seg = pb::qshmm_segments_for(plan.wlen, ae.rho);
          std::vector<uint32_t> seg_state(n_seg + 1, 0);
          bool from_chunks = false;
          if (chain_chunk > 0 && n_seg > 1) {  // k_chain_chunk_err
            const uint32_t n_ch = ae.seg_ok ? (n_seg - 1u + (uint32_t)chain_chunk - 1u) / (uint32_t)chain_chunk : 1u;
            const uint32_t per = n_ch == 1u ? n_seg : (uint32_t)chain_chunk;
            for (uint32_t c = 0; c < n_ch; ++c) {
              const uint32_t k_from = c * per, k_to = std::min(k_from + per, n_seg - 1u);
              uint32_t st = 0, md = T.init_mod;
              bool pz = true;
              if (k_from == 0) {
                pb::errhmm_state_at(T, K, hpp, (uint32_t)read_id, (uint32_t)pass, k_to * PB_TILE, st, md, pz, seg_state.data());
                continue;
              }
              pb::errhmm_segment_start(T, T.emod + (ae.nstates + 1u), ae.reach, K, hpp, (uint32_t)read_id, (uint32_t)pass,
                                       k_from * PB_TILE, ae.seg_ok ? ae.seg_ok : 512u, st, md, pz);
              if (pz) pb::errhmm_state_at(T, K, hpp, (uint32_t)read_id, (uint32_t)pass, k_to * PB_TILE, st, md, pz, seg_state.data());
              else pb::errhmm_chain_range(T, K, (uint32_t)read_id, (uint32_t)pass, st, md, k_from * PB_TILE, k_to * PB_TILE, seg_state.data());
            }
            from_chunks = true;
          } else if (!ae.seg_ok) {
            pb::errhmm_chain_only(T, K, hpp, (uint32_t)read_id, (uint32_t)pass, n_seg, seg_state.data());
          }
          std::vector<uint8_t> slots((size_t)n_seg * PB_TILE + 16, 0);
          std::vector<pb::SegResult> seg(n_seg);
          bool couple_fail = false;
          for (uint32_t k = 0; k < n_seg; ++k) {
            uint32_t st = k == 0 ? 0u : (seg_state[k] & 63u), md = k == 0 ? T.init_mod : ((seg_state[k] >> 6) & 0x3FFu);
            bool pz = k == 0 ? true : (seg_state[k] >> 31) != 0;
            if (k > 0 && ae.seg_ok && !from_chunks) {
              pb::errhmm_segment_start(T, T.emod + (ae.nstates + 1u), ae.reach, K, hpp, (uint32_t)read_id, (uint32_t)pass,
                                       k * PB_TILE, ae.seg_ok, st, md, pz);
            }
            pb::errhmm_simulate_segment(T, K, (uint32_t)read_id, (uint32_t)pass, k * PB_TILE, pz, st, md,
                                        slots.data() + (size_t)k * PB_TILE, seg[k]);
          }
          std::vector<pb::Ckpt> ck(n_seg);
          pb::SegRead sr;
          pb::errhmm_finish_segmented(slots.data(), seg.data(), n_seg, plan.wlen, hpp, K, (uint32_t)read_id, (uint32_t)pass,
                                      ck.data(), sr);
          ++seg_reads;
          if (sr.flags == 0 && !couple_fail) {
            seg_done = true;
            g_out.events.resize(ev_off + sr.ncol);
            memcpy(g_out.events.data() + ev_off, slots.data(), sr.ncol);
            if (hpp.enabled) {  // pass 2's re-derivation of the 4-way choice on non-ACGT bases
              uint32_t R = 0;
              for (uint32_t i = 0; i < sr.ncol; ++i) {
                uint8_t &v = g_out.events[ev_off + i];
                const uint32_t kind = v & 3u;
                if (kind == PB_KIND_SUB && win.nonacgt(R)) {
                  uint32_t w[4];
                  pb::philox_block_keys(K, i, (uint32_t)pass << 16, (uint32_t)read_id, 1u, w);
                  v = (uint8_t)(kind | (((w[0] >> 12) & 3u) << 2));
                }
                R += kind == PB_KIND_INS ? 0u : 1u;
              }
            }
            res.n_entries = sr.ncol; res.rlen = sr.rlen; res.ncol = sr.ncol; res.nsub = sr.nsub; res.nins = sr.nins;
            res.ndel = sr.ndel; res.overflow = 0; res.accuracy = sr.accuracy;
          } else {
            ++seg_fallbacks;
          }
        }
        if (seg_done) {
        } else if (rng_mode == PBSIM_RNG_PHILOX) { pd.pass = (uint32_t)pass; pd.cidx = 0xFFFFFFFFu; pb::errhmm_simulate(T, pd, win, slow, plan.wlen, sink, res); }
        else { rd.cur = cursor; pb::errhmm_simulate(T, rd, win, slow, plan.wlen, sink, res); cursor = rd.cur; }
        if (!seg_done) g_out.events.resize(ev_off + (size_t)res.n_entries);
      }
      if (!(m->method == PBSIM_METHOD_QSHMM && g_out.ckpts.size() != ck_base + cap / PB_TILE + 2))
        g_out.ckpts.resize(ck_base + (res.n_entries + PB_TILE - 1) / PB_TILE);
      int64_t rec[12] = {read_id, pass, (int64_t)plan.acc, (int64_t)plan.offset, (int64_t)plan.wlen, (int64_t)res.rlen,
                         (int64_t)res.ncol, (int64_t)minus, (int64_t)res.n_entries, (int64_t)ev_off, draw_start,
                         (int64_t)res.overflow};
      g_out.info.insert(g_out.info.end(), rec, rec + 12);
      g_out.ck_off.push_back((int64_t)ck_base);
      g_out.counts.push_back(res.nsub); g_out.counts.push_back(res.nins); g_out.counts.push_back(res.ndel);
      g_out.accuracy.push_back(res.accuracy);
      if (pass == 0) len_total += res.rlen;
    }
  }
  return (long)(g_out.info.size() / 12);
}

// --method sample: the engine's schedule (sample_plan.hpp: groups of copies, batches, the quota cut in read order)
// and per-read core (sample_simulate), run sequentially.  Replay mode consumes the log in order, which is the order
// the groups are walked in.
long hostsim_run_sample(const pbsim_model *m, const uint8_t *ascii_upper, const int16_t *hp, long glen, int seq_num,
                        const double bias[12], int rng_mode, uint32_t seed, const int32_t *draws, long ndraws,
                        long long len_quota, long n, const uint8_t *quals, const int64_t *qstart, long batch_reads) {
  pb::ModelImage img;
  if (!img.build(*m)) { fprintf(stderr, "hostsim: %s\n", img.error.c_str()); return -1; }
  img.apply_bias(*m, bias);
  std::vector<uint8_t> hp4((size_t)glen / 2 + 2, 0);
  for (long i = 0; i < glen; ++i) hp4[i >> 1] |= (uint8_t)((hp[i] & 15) << ((i & 1) * 4));
  std::vector<uint8_t> exc((size_t)glen, 0);
  for (long i = 0; i < glen; ++i) {
    const uint8_t c = ascii_upper[i];
    exc[i] = !(c == 'A' || c == 'C' || c == 'G' || c == 'T') || (bias[hp[i] & 15] != 1.0);
  }
  std::vector<uint32_t> xm(((size_t)glen >> 15) + 2, 0);  // 1 bit per 1024-base block with an exceptional base
  for (long i = 0; i < glen; ++i)
    if (exc[i]) xm[(i >> 10) >> 5] |= 1u << ((i >> 10) & 31);
  uint8_t bias_one[12];
  for (int h = 0; h < 12; ++h) bias_one[h] = (bias[h] == 1.0) ? 1 : 0;
  bias_one[0] = 1;
  g_out = HostSimOut();
  pb::QsView T;
  std::memset(&T, 0, sizeof T);
  T.thr = reinterpret_cast<const pb::QsThr *>(img.qs_thr.data());
  T.thr_hp = img.qs_thr_hp.data();
  T.qc_prob = m->qc_prob;
  T.thr32 = reinterpret_cast<const pb::QsThr *>(img.qs_thr32.data());
  T.thr_hp32 = img.qs_thr_hp32.data();
  T.fast = reinterpret_cast<const pb::QsFast *>(img.qs_fast.data());
  pb::SampleSchedule S;
  if (!S.init(len_quota, n, qstart)) return -3;
  long long len_total = 0;
  long reads_done = 0;
  int64_t cursor = 0;
  while (len_total < len_quota) {
    bool first_of_pass = false;
    if (!S.pass_open) {
      if (rng_mode == PBSIM_RNG_PHILOX) S.open_pass(pb::SampleSchedule::philox_value(seed, (uint32_t)seq_num, S.pass, (uint32_t)n));
      else S.open_pass((cursor < ndraws ? draws[cursor] : 0) % n);
      first_of_pass = true;
    }
    pb::SampleGroups G;
    pb::sample_collect(S, batch_reads > 0 ? batch_reads : 1000, 1ll << 40, 1ll << 40, &G);
    const size_t info0 = g_out.info.size() / 12;
    for (size_t g = 0; g < G.entry.size(); ++g) {
      const uint32_t j = G.entry[g];
      uint32_t len = (uint32_t)(qstart[j + 1] - qstart[j]);
      for (uint32_t i = G.first[g]; i < G.first[g + 1]; ++i) {
        const long read_id = reads_done + 1 + i;
        pb::PhiloxDrawQ pd;
        pb::ReplayDraw rd;
        const int64_t draw_start = cursor;
        uint32_t offset = 0;
        if (rng_mode == PBSIM_RNG_PHILOX) {
          pd.ph.k0 = seed; pd.ph.k1 = (uint32_t)seq_num; pd.read_id = (uint32_t)read_id; pd.pass = 0;
          pd.plan_begin();
          if (len >= (uint32_t)glen) len = (uint32_t)glen; else offset = pd.plan_off((uint32_t)glen - len + 1u);
        } else {
          rd.log = draws; rd.cur = cursor + ((first_of_pass && i == 0) ? 1 : 0); rd.end = ndraws; rd.start = cursor;
          if (len >= (uint32_t)glen) len = (uint32_t)glen; else offset = rd.plan_off((uint32_t)glen - len + 1u);
        }
        const uint32_t minus = (read_id % 2 == 0);
        pb::WindowRef win;
        win.ascii = ascii_upper; win.hp4 = hp4.data(); win.offset = offset; win.wlen = len; win.minus = minus;
        bool slow = !img.uniform_bias;
        for (uint32_t k = 0; k < len && !slow; ++k) slow = exc[offset + k];
        const uint32_t cap = len * 2 + 4096;
        const size_t ck_base = g_out.ckpts.size();
        g_out.ckpts.resize(ck_base + cap / PB_TILE + 2);
        size_t ev_off = (g_out.events.size() + 15) / 16 * 16;
        g_out.events.resize(ev_off + (size_t)cap * 2);
        pb::QsSink sink;
        sink.init(reinterpret_cast<uint16_t *>(g_out.events.data() + ev_off), g_out.ckpts.data() + ck_base, cap);
        pb::SubreadResult res;
        bool seg_done = false;
        if (rng_mode == PBSIM_RNG_PHILOX && use_seg && img.uniform_bias && (int)len >= seg_min_len) {
          // (the engine segments a copy only when it assumed the right length; here every copy's length is known)
          pb::HpProbe hpp;
          hpp.enabled = slow ? 1u : 0u;
          hpp.win = win;
          hpp.xm = xm.data();
          hpp.bias_one = bias_one;
          ++seg_reads;
          seg_done = run_segmented_sample(T, seed, (uint32_t)seq_num, (uint32_t)read_id, len, quals + qstart[j], hpp,
                                          g_out.events, ev_off, g_out.ckpts, ck_base, res);
          if (!seg_done) ++seg_fallbacks;
        }
        if (seg_done) {
        } else if (rng_mode == PBSIM_RNG_PHILOX) {
          g_out.events.resize(ev_off + (size_t)cap * 2);
          g_out.ckpts.resize(ck_base + cap / PB_TILE + 2);
          sink.init(reinterpret_cast<uint16_t *>(g_out.events.data() + ev_off), g_out.ckpts.data() + ck_base, cap);
          pb::sample_simulate(T, pd, win, slow, len, quals + qstart[j], sink, res);
        } else { pb::sample_simulate(T, rd, win, slow, len, quals + qstart[j], sink, res); cursor = rd.cur; }
        if (!seg_done) {
          g_out.events.resize(ev_off + (size_t)res.n_entries * 2);
          g_out.ckpts.resize(ck_base + (res.n_entries + PB_TILE - 1) / PB_TILE);
        }
        int64_t rec[12] = {read_id, 0, 0, (int64_t)offset, (int64_t)len, (int64_t)res.rlen, (int64_t)res.ncol,
                           (int64_t)minus, (int64_t)res.n_entries, (int64_t)ev_off, draw_start, (int64_t)res.overflow};
        g_out.info.insert(g_out.info.end(), rec, rec + 12);
        g_out.ck_off.push_back((int64_t)ck_base);
        g_out.counts.push_back(res.nsub); g_out.counts.push_back(res.nins); g_out.counts.push_back(res.ndel);
        g_out.accuracy.push_back(res.accuracy);
        len = res.rlen;  // the next copy is as long as this read (:1835, :1756)
      }
    }
    // the quota test in front of every read (:1747, :1753), in read order
    const size_t nb = g_out.info.size() / 12 - info0;
    size_t keep = nb;
    for (size_t r = 0; r < nb; ++r) {
      if (len_total >= len_quota) { keep = r; break; }
      len_total += g_out.info[(info0 + r) * 12 + 5];
    }
    if (keep < nb) {
      g_out.info.resize((info0 + keep) * 12);
      g_out.counts.resize((info0 + keep) * 3);
      g_out.accuracy.resize(info0 + keep);
      g_out.ck_off.resize(info0 + keep);
      break;
    }
    reads_done += (long)nb;
    S.j = G.j_end;
    if (S.j >= S.n) S.next_pass();
  }
  return (long)(g_out.info.size() / 12);
}

// the planners of --strategy trans / templ (plan_read_trans / plan_read_templ, the functions k_plan calls) for every
// read of a set in PHILOX mode; out: 4 per read (offset inside the sequence, window length, accuracy, minus)
long hostsim_plan_set(const pbsim_model *m, int strategy, long n, const int64_t *tlen, const int32_t *plus,
                      const int32_t *minus, uint32_t seed, int rank_max, int64_t *out, long cap_reads) {
  std::vector<uint16_t> ends((size_t)(rank_max + 1) * 21), mod((size_t)rank_max + 1);
  if (strategy == PBSIM_STRATEGY_TRANS) pbsim_host_ssp_table(rank_max, ends.data(), mod.data());
  pb::PlanTables pt;
  pt.prob2len = m->prob2len;
  pt.prob2acc = m->prob2accuracy;
  pt.len_rand_value = (uint32_t)m->len_rand_value;
  pt.acc_rand_value = (uint32_t)m->accuracy_rand_value;
  pt.len_min = (uint32_t)m->len_min;
  long read_id = 0;
  for (long t = 0; t < n; ++t) {
    const long nr = strategy == PBSIM_STRATEGY_TRANS ? (long)plus[t] + (long)minus[t] : 1;
    for (long k = 1; k <= nr; ++k) {
      ++read_id;
      if (read_id > cap_reads) return -1;
      pb::PhiloxDraw d;
      d.ph.k0 = seed; d.ph.k1 = 0u; d.read_id = (uint32_t)read_id; d.pass = 0;
      const pb::ReadPlan p = strategy == PBSIM_STRATEGY_TRANS
                                 ? pb::plan_read_trans(pt, d, ends.data(), mod.data(), (uint32_t)tlen[t])
                                 : pb::plan_read_templ(pt, d, (uint32_t)tlen[t]);
      int64_t *o = out + (read_id - 1) * 4;
      o[0] = p.offset; o[1] = p.wlen; o[2] = p.acc;
      o[3] = strategy == PBSIM_STRATEGY_TRANS ? (k <= plus[t] ? 0 : 1) : 0;
    }
  }
  return read_id;
}

const int64_t *hostsim_info() { return g_out.info.data(); }
const uint32_t *hostsim_counts() { return g_out.counts.data(); }
const double *hostsim_accuracy() { return g_out.accuracy.data(); }
const uint8_t *hostsim_events(long *nbytes) { *nbytes = (long)g_out.events.size(); return g_out.events.data(); }
const pb::Ckpt *hostsim_ckpts(long *n) { *n = (long)g_out.ckpts.size(); return g_out.ckpts.data(); }
const int64_t *hostsim_ck_off() { return g_out.ck_off.data(); }
long hostsim_draws_consumed_total() { return 0; }

}  // extern "C"
