// --method sample: which pool entry every read copies, restated from the sequential loop of
// simulate_by_sample (pbsim.cpp:1717-1770) as an index computation.
//
// The pool is what get_sample_inf (:1214-1275) leaves in fp_filtered: the quality strings that pass the length
// and accuracy filters, in file order.  The reference walks the pool again and again ("pool passes") until the
// quota is met:
//   pass 0      : entry j is copied sample_num times, plus once more if (sample_value + j) % sample_interval == 0
//   pass p >= 1 : sample_num = 0, only the extra copy
// with sample_value = rand() % num_filtered drawn at the start of every pass (:1734).  The copies of one entry
// follow each other (consecutive read numbers) and form a CHAIN: the quality buffer is cut where a read ended
// (mut.qc[read_offset] = 0, :1835) and measured again for the next copy (mut.len = strlen(mut.qc), :1756), so copy
// i+1 is as long as copy i's read.  The engine therefore runs one GPU thread per (entry, pass) GROUP and the groups
// of a pass in parallel; this header only enumerates the groups.  Host code, shared by the engine and tests/hostsim.
#pragma once
#include <stdint.h>

#include <vector>

#include "sim_core.cuh"

namespace pb {

struct SampleSchedule {
  int64_t n = 0;                  // sample.num_filtered
  const int64_t *qstart = nullptr;  // [n+1] offsets of the quality strings
  int64_t num0 = 0;               // sample_num of pass 0 (:1719)
  int64_t interval = 1;           // sample_interval (:1720-1730)
  uint32_t pass = 0;              // pool pass being walked
  int64_t value0 = 0;             // sample_value at entry 0 of this pass
  int64_t j = 0;                  // next entry of this pass
  bool pass_open = false;         // value0 is known for `pass`

  // false: the reference divides by zero (:1723) or has nothing to copy
  bool init(int64_t len_quota, int64_t n_, const int64_t *qstart_) {
    n = n_;
    qstart = qstart_;
    if (n < 2 || qstart[n] < 1) return false;
    const int64_t pool_total = qstart[n];
    num0 = len_quota / pool_total;
    const int64_t residue = len_quota % pool_total;
    if (residue == 0) {
      interval = 1;
    } else {
      interval = (int64_t)((double)(pool_total / residue) * 2 + 0.5);
      if (interval > (int64_t)(n * 0.5)) interval = (int64_t)(n * 0.5);
    }
    if (interval < 1) return false;
    pass = 0;
    j = 0;
    pass_open = false;
    return true;
  }
  // the pool-pass draw in PHILOX mode: domain 3, counter = pass
  static uint32_t philox_value(uint32_t seed, uint32_t seq_num, uint32_t pass, uint32_t n) {
    Philox ph;
    ph.k0 = seed;
    ph.k1 = seq_num;
    uint32_t w[4];
    ph.block(pass, 0u, 0u, 3u, w);
    return mulhi32(w[0], n);
  }
  void open_pass(int64_t sample_value) {
    value0 = sample_value;
    pass_open = true;
  }
  int64_t copies(int64_t entry) const {
    return (pass == 0 ? num0 : 0) + (((value0 + entry) % interval) == 0 ? 1 : 0);
  }
  void next_pass() {
    ++pass;
    j = 0;
    pass_open = false;
  }
};

// one batch: groups of the current pass, from entry S.j on, until `max_reads` reads or `max_bases` planned bases
struct SampleGroups {
  std::vector<uint32_t> entry;  // pool entry of group g
  std::vector<uint32_t> first;  // [G+1] first read of group g inside the batch
  int64_t j_end = 0;            // first entry after the batch
};

// hard_reads: never plan more reads than this (replay: the reads the log holds; the last group may be cut short,
// its missing copies lie behind the quota cut)
inline void sample_collect(const SampleSchedule &S, int64_t max_reads, int64_t max_bases, int64_t hard_reads,
                           SampleGroups *out) {
  out->entry.clear();
  out->first.clear();
  int64_t reads = 0, bases = 0, j = S.j;
  for (; j < S.n; ++j) {
    int64_t c = S.copies(j);
    if (c == 0) continue;
    if (reads > 0 && (reads + c > max_reads || bases >= max_bases)) break;
    if (reads + c > hard_reads) c = hard_reads - reads;
    if (c <= 0) break;
    out->entry.push_back((uint32_t)j);
    out->first.push_back((uint32_t)reads);
    reads += c;
    bases += c * (S.qstart[j + 1] - S.qstart[j]);
  }
  out->first.push_back((uint32_t)reads);
  out->j_end = j;
}

}  // namespace pb
