// Host half of the GPU gzip writer (gz_kernels.cuh): one Huffman code per record stream and batch.
//
// The reference pipes its records through `popen("gzip > file")` (pbsim.cpp:708-730, :771-788); its outputs are
// gzip files.  With option "deflate" the engine hands out gzip MEMBERS instead of text: every 32 KiB of a record
// stream becomes one member holding a single dynamic-Huffman DEFLATE block of literals (no LZ77 matches: the
// streams are near-random 4-letter text and quality strings, where matches buy little), encoded on the device.
// All members of a batch share one code, built here from the stream's byte histogram:
//   * code lengths by Huffman's algorithm, limited to kGzMaxBits by halving the counts until the tree fits
//     (every literal keeps a non-zero count, so any byte can be coded);
//   * canonical codes as RFC 1951 3.2.2 assigns them, stored bit-reversed (DEFLATE packs codes MSB first into an
//     LSB-first bit stream);
//   * the block header (BFINAL = 1, BTYPE = 2, HLIT = 257, HDIST = 2, code lengths run-length coded with their own
//     Huffman code, RFC 1951 3.2.7) as a ready-made bit string.
// Concatenated members decompress to exactly the text the engine would have delivered.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <queue>
#include <utility>
#include <vector>

namespace pb {

constexpr uint32_t kGzMaxBits = 12;   // longest literal code (bounds a member's size: 1.5 bytes per input byte)
constexpr uint32_t kGzHdrWords = 80;  // room for the dynamic block header (at most ~2.4 kbit)

struct GzCode {
  uint32_t lit[256];  // bit-reversed code | length << 16
  uint32_t eob;       // same packing, end-of-block symbol 256
  uint32_t hdr_bits;  // block header: 3 bits + dynamic header
  uint32_t hdr[kGzHdrWords];
  uint8_t len[256];   // code lengths alone (size pass)
};

// code lengths (0 for unused symbols) of a Huffman code over freq[], none longer than limit
inline std::vector<uint8_t> gz_code_lengths(std::vector<uint64_t> freq, uint32_t limit) {
  const size_t n = freq.size();
  std::vector<uint8_t> len(n, 0);
  size_t used = 0;
  for (uint64_t f : freq) used += f != 0;
  if (used == 0) return len;
  if (used == 1) {  // a one-symbol code still needs one bit (and a complete code needs a partner)
    for (size_t i = 0; i < n; ++i)
      if (freq[i]) len[i] = 1;
    for (size_t i = 0; i < n; ++i)
      if (!freq[i]) {
        len[i] = 1;
        break;
      }
    return len;
  }
  for (;;) {
    // nodes: leaves 0..n-1, internal n..; parent links give the depths
    std::vector<uint64_t> w(freq);
    std::vector<int> parent(2 * n, -1);
    typedef std::pair<uint64_t, int> item;  // (weight, node); ties broken by node index: deterministic
    std::priority_queue<item, std::vector<item>, std::greater<item>> pq;
    for (size_t i = 0; i < n; ++i)
      if (freq[i]) pq.push(item(freq[i], (int)i));
    int next = (int)n;
    while (pq.size() > 1) {
      const item a = pq.top();
      pq.pop();
      const item b = pq.top();
      pq.pop();
      parent[a.second] = next;
      parent[b.second] = next;
      pq.push(item(a.first + b.first, next));
      ++next;
    }
    uint32_t maxlen = 0;
    for (size_t i = 0; i < n; ++i) {
      if (!freq[i]) continue;
      uint32_t d = 0;
      for (int v = (int)i; parent[v] >= 0; v = parent[v]) ++d;
      len[i] = (uint8_t)d;
      maxlen = std::max(maxlen, d);
    }
    if (maxlen <= limit) return len;
    for (uint64_t &f : freq)
      if (f) f = (f + 1) >> 1;  // flatten the distribution and try again
  }
}

// canonical codes (RFC 1951 3.2.2), bit-reversed over their length
inline std::vector<uint32_t> gz_canonical_reversed(const std::vector<uint8_t> &len) {
  uint32_t bl_count[17] = {0}, next_code[17] = {0};
  for (uint8_t l : len) bl_count[l]++;
  bl_count[0] = 0;
  uint32_t code = 0;
  for (int bits = 1; bits <= 16; ++bits) {
    code = (code + bl_count[bits - 1]) << 1;
    next_code[bits] = code;
  }
  std::vector<uint32_t> out(len.size(), 0);
  for (size_t i = 0; i < len.size(); ++i) {
    const uint32_t l = len[i];
    if (!l) continue;
    uint32_t c = next_code[l]++, r = 0;
    for (uint32_t b = 0; b < l; ++b) r |= ((c >> b) & 1u) << (l - 1 - b);
    out[i] = r;
  }
  return out;
}

struct GzBitWriter {
  uint32_t *w;
  uint32_t cap_words, bits = 0;
  void put(uint32_t value, uint32_t n) {  // n <= 16, LSB first
    for (uint32_t b = 0; b < n; ++b, ++bits)
      if (bits < cap_words * 32u && ((value >> b) & 1u)) w[bits >> 5] |= 1u << (bits & 31u);
  }
};

// hist: byte counts of the stream (any byte may be zero: it still gets a code)
inline void gz_build_code(const uint64_t hist[256], GzCode *out, bool flat = false) {
  std::vector<uint64_t> freq(257);
  for (int i = 0; i < 256; ++i) freq[i] = (hist[i] && !flat) ? hist[i] : 1;
  freq[256] = 1;  // end of block: once per member
  const std::vector<uint8_t> ll = gz_code_lengths(freq, kGzMaxBits);
  const std::vector<uint32_t> lc = gz_canonical_reversed(ll);
  for (int i = 0; i < 256; ++i) {
    out->lit[i] = lc[i] | ((uint32_t)ll[i] << 16);
    out->len[i] = ll[i];
  }
  out->eob = lc[256] | ((uint32_t)ll[256] << 16);

  // ---- block header
  // distance alphabet: no distance is ever used; like zlib, send two 1-bit codes so that the code is complete
  std::vector<uint8_t> seq(ll.begin(), ll.end());  // 257 literal/length code lengths (HLIT = 257)
  seq.push_back(1);
  seq.push_back(1);  // HDIST = 2
  // run-length code the sequence (RFC 1951 3.2.7): 16 = repeat previous 3-6, 17 = zeros 3-10, 18 = zeros 11-138
  struct Tok { uint8_t sym, extra_bits; uint16_t extra; };
  std::vector<Tok> toks;
  for (size_t i = 0; i < seq.size();) {
    size_t j = i;
    while (j < seq.size() && seq[j] == seq[i]) ++j;
    size_t run = j - i;
    if (seq[i] == 0) {
      while (run >= 11) {
        const size_t r = std::min<size_t>(run, 138);
        toks.push_back({18, 7, (uint16_t)(r - 11)});
        run -= r;
      }
      if (run >= 3) {
        toks.push_back({17, 3, (uint16_t)(run - 3)});
        run = 0;
      }
      while (run--) toks.push_back({0, 0, 0});
    } else {
      toks.push_back({seq[i], 0, 0});
      --run;
      while (run >= 3) {
        const size_t r = std::min<size_t>(run, 6);
        toks.push_back({16, 2, (uint16_t)(r - 3)});
        run -= r;
      }
      while (run--) toks.push_back({seq[i], 0, 0});
    }
    i = j;
  }
  std::vector<uint64_t> clf(19, 0);
  for (const Tok &t : toks) clf[t.sym]++;
  const std::vector<uint8_t> cll = gz_code_lengths(clf, 7);
  const std::vector<uint32_t> clc = gz_canonical_reversed(cll);
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int hclen = 19;
  while (hclen > 4 && cll[order[hclen - 1]] == 0) --hclen;
  for (uint32_t &x : out->hdr) x = 0;
  GzBitWriter bw{out->hdr, kGzHdrWords};
  bw.put(1, 1);  // BFINAL
  bw.put(2, 2);  // BTYPE = dynamic Huffman
  bw.put(257 - 257, 5);
  bw.put(2 - 1, 5);
  bw.put((uint32_t)(hclen - 4), 4);
  for (int i = 0; i < hclen; ++i) bw.put(cll[order[i]], 3);
  for (const Tok &t : toks) {
    bw.put(clc[t.sym], cll[t.sym]);
    if (t.extra_bits) bw.put(t.extra, t.extra_bits);
  }
  out->hdr_bits = bw.bits;
  // a header that does not fit its buffer (never seen; a few hundred distinct code lengths in adversarial order)
  // falls back to the flat 8/9-bit code, whose header is a handful of run codes
  if (bw.bits > kGzHdrWords * 32u && !flat) gz_build_code(hist, out, true);
}

// x^(2^k) mod P of the reflected CRC-32 polynomial, k = 0..31 (the table zlib's crc32_combine builds)
inline uint32_t gz_gf_mul(uint32_t a, uint32_t b) {
  uint32_t m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1)) == 0) break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
  }
  return p;
}
inline void gz_x2n_table(uint32_t t[32]) {
  uint32_t p = 1u << 30;  // x^1
  t[0] = p;
  for (int n = 1; n < 32; ++n) t[n] = p = gz_gf_mul(p, p);
}
// byte-wise CRC-32 table (reflected, polynomial 0xEDB88320)
inline void gz_crc_table(uint32_t t[256]) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
    t[i] = c;
  }
}

}  // namespace pb
