// GPU gzip writer — device half (host half: gz_host.hpp).  Replaces popen("gzip > file") (pbsim.cpp:708-730).
//
// A record stream is cut into UNITS of kGzUnit input bytes; every unit becomes one gzip member:
//   header (10 bytes, FNAME flag) | name: k padding characters + NUL | DEFLATE blocks | CRC-32 | ISIZE
// k in 0..3 makes the member's size a multiple of 4, so every member starts 4-byte aligned in the output and a
// member built in shared memory is copied out with whole-word, coalesced stores.
// A member holds up to kGzBlocks dynamic-Huffman blocks of literals, one per 4 KiB of text (one warp's slices), and
// every block is coded with the cheapest of THREE codes of the stream: code 0 is built from the text that looks like
// sequence (FASTQ sequence lines, MAF rows: A C G T -), code 1 from everything else (quality lines, headers), code 2
// from all of it (blocks that mix both, as with short reads).  A record stream interleaves 2-bit and 5-bit material
// line by line; one code for both costs 10-25 % of the ratio.
//   k_gz_hist   byte histograms of the two classes (the codes are built from them on the host)
//   k_gz_size   per unit: bit counts of every 128-byte slice under both codes, the block's choice, the exact
//               member size and the CRC-32 of its text
//   (scan)      member offsets
//   k_gz_encode per unit: one CTA, 256 threads x 128-byte slices; scans of the recorded bit counts give every thread
//               the bit position of its codes, which it ORs into the shared-memory image
// Both passes read the text once (4.1 B per emitted base each); the encoder writes about a third of that.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gz_host.hpp"

namespace pb {

constexpr uint32_t kGzUnit = 32768;                        // input bytes per member
constexpr uint32_t kGzThreads = 256;
constexpr uint32_t kGzSlice = kGzUnit / kGzThreads;        // 128 bytes per thread
constexpr uint32_t kGzBlocks = kGzThreads / 32;            // DEFLATE blocks per member: one per warp (4 KiB of text)
constexpr uint32_t kGzImgBytes = 24 + kGzBlocks * (kGzHdrWords * 4 + 4) + kGzUnit / 8 * kGzMaxBits + 16;  // worst case
constexpr uint32_t kGzImgWords = (kGzImgBytes + 3) / 4;

constexpr uint32_t kGzCodes = 3;
struct GzTables {        // device copy of the GzCodes + CRC constants
  unsigned long long len3[256];  // lengths under codes 0, 1, 2 in 16-bit fields (size pass: all at once)
  uint32_t lit[kGzCodes][256];   // bit-reversed code | length << 16
  uint32_t crc[256];
  uint32_t hdr[kGzCodes][kGzHdrWords];
  uint32_t x2n[32];      // x^(2^k) mod P
  uint32_t tail[kGzThreads];  // x^(8 * kGzSlice * (kGzThreads-1-t)) mod P: moves slice t's CRC to the end of a full unit
  uint32_t eob[kGzCodes], hdr_bits[kGzCodes];
  uint32_t pad[2];
};

// text that looks like sequence: at least 13 of 16 bytes in "ACGT-" (quality strings hold those letters too, but
// never that densely)
__device__ __forceinline__ uint32_t gz_is_seq_byte(uint32_t c) {
  return (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == '-') ? 1u : 0u;
}

__global__ void k_gz_hist(const uint8_t *__restrict__ in, uint64_t n, uint32_t sample, unsigned long long *hist /*[2][256]*/) {
  __shared__ unsigned int h[2][256];
  for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) (&h[0][0])[i] = 0;
  __syncthreads();
  // a sample is enough (every byte value gets a code whatever the counts): 16 bytes per thread out of every
  // `sample` chunks, spread over the whole stream
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16u * sample;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16u * sample; i < n; i += stride) {
    if (i + 16u <= n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + i));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t nseq = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int b = 0; b < 4; ++b) nseq += gz_is_seq_byte((w[q] >> (8 * b)) & 0xFFu);
      unsigned int *hh = h[nseq >= 13u ? 0 : 1];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int b = 0; b < 4; ++b) atomicAdd(&hh[(w[q] >> (8 * b)) & 0xFFu], 1u);
    } else {
      for (uint64_t j = i; j < n; ++j) atomicAdd(&h[1][in[j]], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x)
    if ((&h[0][0])[i]) atomicAdd(&hist[i], (unsigned long long)(&h[0][0])[i]);
}

// a * b mod P in the reflected representation zlib's crc32_combine uses (bit 31 = x^0); branch-free, 32 steps
__device__ __forceinline__ uint32_t gz_gf_mul_dev(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    p ^= (0u - (a >> 31)) & b;
    a <<= 1;
    b = (b >> 1) ^ ((0u - (b & 1u)) & 0xEDB88320u);
  }
  return p;
}
// x^(8 n) mod P
__device__ __forceinline__ uint32_t gz_x8n(const uint32_t *x2n, uint64_t n) {
  uint32_t p = 1u << 31;  // x^0
  uint32_t k = 3;
  while (n) {
    if (n & 1u) p = gz_gf_mul_dev(x2n[k & 31u], p);
    n >>= 1;
    ++k;
  }
  return p;
}

// member geometry from the body's bit count.  bgzf: the 18-byte BGZF header (FEXTRA with the 'BC' subfield holding
// the block size, SAM spec 4.1; htslib insists on XLEN == 6, so no padding and no alignment), else the 10-byte header
// with a 0-3 character FNAME that makes the member's size a multiple of 4
__device__ __forceinline__ void gz_geometry(uint32_t body_bits, uint32_t bgzf, uint32_t *body_bytes, uint32_t *k,
                                            uint32_t *hdr_bytes, uint32_t *total) {
  const uint32_t B = (body_bits + 7u) >> 3;
  *body_bytes = B;
  if (bgzf) {
    *k = 0;
    *hdr_bytes = 18u;
    *total = 18u + B + 8u;
  } else {
    const uint32_t kk = (4u - ((19u + B) & 3u)) & 3u;
    *k = kk;
    *hdr_bytes = 11u + kk;    // 10 header + k name + NUL
    *total = 19u + kk + B;    // ... + body + 8 trailer
  }
}

// CTA-wide sums / scans over kGzThreads values
__device__ __forceinline__ uint32_t gz_block_excl_scan(uint32_t v, uint32_t *warp_tot /*[8] smem*/, uint32_t *total) {
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31u) warp_tot[w] = x;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (uint32_t q = 0; q < kGzThreads / 32u; ++q) {
    const uint32_t t = warp_tot[q];
    if (q < w) base += t;
    tot += t;
  }
  *total = tot;
  __syncthreads();
  return base + x - v;
}

// per-thread: bit counts under the three codes (16-bit fields) and CRC-32 of its slice
__device__ __forceinline__ void gz_slice_scan(const uint8_t *__restrict__ in, uint64_t pos, uint32_t nbytes,
                                              const unsigned long long *len2_s, const uint32_t *crc_s,
                                              unsigned long long *bits2, uint32_t *crc_out) {
  unsigned long long nb = 0;
  uint32_t crc = 0xFFFFFFFFu;
  if (nbytes == kGzSlice) {
#pragma unroll 2
    for (uint32_t q = 0; q < kGzSlice; q += 16u) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + pos + q));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = (w[a] >> (8 * b)) & 0xFFu;
          nb += len2_s[c];
          crc = crc_s[(crc ^ c) & 0xFFu] ^ (crc >> 8);
        }
    }
  } else {
    for (uint32_t q = 0; q < nbytes; ++q) {
      const uint32_t c = in[pos + q];
      nb += len2_s[c];
      crc = crc_s[(crc ^ c) & 0xFFu] ^ (crc >> 8);
    }
  }
  *bits2 = nb;
  *crc_out = nbytes ? ~crc : 0u;  // finalised CRC of the slice (0 for an empty slice)
}

// unit_size[u] = member bytes, unit_crc[u] = CRC-32 of the unit's text, unit_sel[u] = code of each of its blocks
// (bits 2w, 2w+1: the code of block w), slice_bits[u * kGzThreads + t] = bits of slice t under its block's code.
// The input is padded to a multiple of 16.
__global__ void __launch_bounds__(kGzThreads) k_gz_size(const uint8_t *__restrict__ in, uint64_t n, const GzTables *T,
                                                        uint32_t bgzf, uint32_t *unit_size, uint32_t *unit_crc,
                                                        uint16_t *unit_sel, uint16_t *slice_bits) {
  __shared__ unsigned long long len2_s[256];
  __shared__ uint32_t crc_s[256];
  __shared__ uint32_t red[kGzThreads / 32];
  __shared__ uint32_t blk_bits[kGzBlocks], blk_sel[kGzBlocks];
  len2_s[threadIdx.x] = T->len3[threadIdx.x];
  crc_s[threadIdx.x] = T->crc[threadIdx.x];
  __syncthreads();
  const uint64_t u0 = (uint64_t)blockIdx.x * kGzUnit;
  const uint32_t un = (uint32_t)min((uint64_t)kGzUnit, n - u0);
  const uint32_t s0 = threadIdx.x * kGzSlice;
  const uint32_t sn = s0 >= un ? 0u : min(kGzSlice, un - s0);
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  unsigned long long bits3;
  uint32_t crc;
  gz_slice_scan(in, u0 + s0, sn, len2_s, crc_s, &bits3, &crc);
  // the block (this warp's 4 KiB) takes the cheapest code, headers and end-of-block symbols included
  uint32_t cost[kGzCodes], mine[kGzCodes];
#pragma unroll
  for (uint32_t c = 0; c < kGzCodes; ++c) {
    mine[c] = (uint32_t)(bits3 >> (16u * c)) & 0xFFFFu;
    cost[c] = __reduce_add_sync(0xFFFFFFFFu, mine[c]) + T->hdr_bits[c] + (T->eob[c] >> 16);
  }
  uint32_t sel = 0;
  if (cost[1] < cost[sel]) sel = 1;
  if (cost[2] < cost[sel]) sel = 2;
  const bool nonempty = w * 32u * kGzSlice < un;
  if (lane == 0) {
    blk_bits[w] = nonempty ? cost[sel] : 0u;
    blk_sel[w] = nonempty ? sel : 0u;
  }
  slice_bits[(uint64_t)blockIdx.x * kGzThreads + threadIdx.x] = (uint16_t)(sel == 0u ? mine[0] : (sel == 1u ? mine[1] : mine[2]));
  // crc(A || B) = crc(A) * x^(8 |B|) + crc(B)  =>  crc(unit) = XOR over slices of crc(slice) * x^(8 * bytes after it):
  // one multiplication per thread (a table constant for full units), then an XOR reduction
  const uint32_t after = un - (s0 + sn);  // bytes of the unit behind this slice
  uint32_t part = 0;
  if (sn != 0u) part = gz_gf_mul_dev(un == kGzUnit ? T->tail[threadIdx.x] : gz_x8n(T->x2n, after), crc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, o);
  if (lane == 0u) red[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t c = 0, body = 0, selmask = 0;
    for (uint32_t q = 0; q < kGzBlocks; ++q) {
      c ^= red[q];
      body += blk_bits[q];
      selmask |= blk_sel[q] << (2u * q);
    }
    if (body == 0u) body = T->hdr_bits[0] + (T->eob[0] >> 16);  // an empty unit still holds one (empty) block
    uint32_t B, k, hb, total;
    gz_geometry(body, bgzf, &B, &k, &hb, &total);
    unit_size[blockIdx.x] = total;
    unit_crc[blockIdx.x] = c;
    unit_sel[blockIdx.x] = (uint16_t)selmask;
  }
}

__device__ __forceinline__ void gz_put_bits(uint32_t *img, uint32_t bitpos, uint32_t value, uint32_t nbits) {
  // nbits <= 32; value has no bits above nbits
  const uint32_t w = bitpos >> 5, sh = bitpos & 31u;
  atomicOr(&img[w], value << sh);
  if (sh + nbits > 32u) atomicOr(&img[w + 1u], value >> (32u - sh));
}

__global__ void __launch_bounds__(kGzThreads) k_gz_encode(const uint8_t *__restrict__ in, uint64_t n, const GzTables *T,
                                                          const uint64_t *unit_off, const uint32_t *unit_crc,
                                                          const uint16_t *unit_sel, const uint16_t *slice_bits,
                                                          uint32_t bgzf, uint8_t *out) {
  extern __shared__ __align__(16) uint32_t gz_smem[];
  uint32_t *img = gz_smem;                       // [kGzImgWords]
  uint32_t *lit_s = gz_smem + kGzImgWords;       // [kGzCodes][256]
  __shared__ uint32_t blk_body[kGzBlocks];
  for (uint32_t i = threadIdx.x; i < kGzImgWords; i += kGzThreads) img[i] = 0;
#pragma unroll
  for (uint32_t c = 0; c < kGzCodes; ++c) lit_s[c * 256u + threadIdx.x] = T->lit[c][threadIdx.x];
  const uint64_t u0 = (uint64_t)blockIdx.x * kGzUnit;
  const uint32_t un = (uint32_t)min((uint64_t)kGzUnit, n - u0);
  const uint32_t s0 = threadIdx.x * kGzSlice;
  const uint32_t sn = s0 >= un ? 0u : min(kGzSlice, un - s0);
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t selmask = unit_sel[blockIdx.x];
  const uint32_t sel = (selmask >> (2u * w)) & 3u;
  // the slice's bit count was recorded by the size pass: position inside the block by a warp scan
  const uint32_t bits = slice_bits[(uint64_t)blockIdx.x * kGzThreads + threadIdx.x];
  uint32_t x = bits;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  const uint32_t in_block = x - bits;
  if (lane == 31u) blk_body[w] = x;
  __syncthreads();
  // blocks before this one: header + body + end of block each (empty blocks of a short last unit are left out)
  const uint32_t nblk = un == 0u ? 1u : (un + 32u * kGzSlice - 1u) / (32u * kGzSlice);
  uint32_t blk_pos = 0, total_bits = 0;
  for (uint32_t q = 0; q < nblk; ++q) {
    const uint32_t sq = (selmask >> (2u * q)) & 3u;
    const uint32_t bb = T->hdr_bits[sq] + blk_body[q] + (T->eob[sq] >> 16);
    if (q < w) blk_pos += bb;
    total_bits += bb;
  }
  uint32_t B, k, hb, total;
  gz_geometry(total_bits, bgzf, &B, &k, &hb, &total);
  const uint32_t body0 = hb * 8u;  // bit position of the DEFLATE stream inside the member
  const uint32_t hdr_bits = T->hdr_bits[sel], eob = T->eob[sel];
  // member header and trailer
  if (threadIdx.x == 0) {
    if (bgzf) {
      gz_put_bits(img, 0, 0x04088B1Fu, 32);       // ID1 ID2 CM=8 FLG=FEXTRA
      gz_put_bits(img, 64, 0xFF00u, 16);          // XFL = 0, OS = 255 (unknown), as htslib writes it
      gz_put_bits(img, 80, 6u, 16);               // XLEN
      gz_put_bits(img, 96, 0x00024342u, 32);      // 'B' 'C' SLEN = 2
      gz_put_bits(img, 128, total - 1u, 16);      // BSIZE: block size minus 1
    } else {
      gz_put_bits(img, 0, 0x08088B1Fu, 32);       // ID1 ID2 CM=8 FLG=FNAME
      gz_put_bits(img, 64, 0x0300u, 16);          // XFL = 0, OS = 3 (Unix); MTIME stays 0
      for (uint32_t j = 0; j < k; ++j) gz_put_bits(img, 80u + 8u * j, (uint32_t)'p', 8);
    }
    const uint32_t tr = (hb + B) * 8u;
    gz_put_bits(img, tr, unit_crc[blockIdx.x], 32);
    gz_put_bits(img, tr + 32u, un, 32);           // ISIZE
  }
  if (w < nblk) {
    // block header (the ready-made bit string starts with BFINAL = 1: cleared for every block but the last) and
    // the end-of-block symbol behind the block's codes
    const uint32_t p0 = body0 + blk_pos;
    for (uint32_t i = lane; i * 32u < hdr_bits; i += 32u) {
      const uint32_t nb = min(32u, hdr_bits - i * 32u);
      uint32_t v = nb == 32u ? T->hdr[sel][i] : (T->hdr[sel][i] & ((1u << nb) - 1u));
      if (i == 0u && w + 1u < nblk) v &= ~1u;
      gz_put_bits(img, p0 + i * 32u, v, nb);
    }
    if (lane == 31u) gz_put_bits(img, p0 + hdr_bits + blk_body[w], eob & 0xFFFFu, eob >> 16);
    // the slice's codes
    uint32_t pos = p0 + hdr_bits + in_block;
    uint32_t wd = pos >> 5;
    uint32_t accbits = pos & 31u;
    unsigned long long acc = 0ull;
    const uint32_t *lit = lit_s + sel * 256u;
    auto push = [&](uint32_t c) {
      const uint32_t e = lit[c];
      acc |= (unsigned long long)(e & 0xFFFFu) << accbits;
      accbits += e >> 16;
      if (accbits >= 32u) {
        atomicOr(&img[wd], (uint32_t)acc);
        ++wd;
        acc >>= 32;
        accbits -= 32u;
      }
    };
    if (sn == kGzSlice) {
#pragma unroll 2
      for (uint32_t q = 0; q < kGzSlice; q += 16u) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + u0 + s0 + q));
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) push((ww[a] >> (8 * b)) & 0xFFu);
      }
    } else {
      for (uint32_t q = 0; q < sn; ++q) push(in[u0 + s0 + q]);
    }
    if (accbits) atomicOr(&img[wd], (uint32_t)acc);
  }
  __syncthreads();
  if (bgzf) {  // byte-granular placement
    uint8_t *dst = out + unit_off[blockIdx.x];
    const uint8_t *img8 = reinterpret_cast<const uint8_t *>(img);
    for (uint32_t i = threadIdx.x; i < total; i += kGzThreads) dst[i] = img8[i];
  } else {     // whole words: the member starts 4-byte aligned and its size is a multiple of 4
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + unit_off[blockIdx.x]);
    const uint32_t nw = total >> 2;
    for (uint32_t i = threadIdx.x; i < nw; i += kGzThreads) dst[i] = img[i];
  }
}

}  // namespace pb
