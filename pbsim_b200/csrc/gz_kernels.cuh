// GPU gzip writer — device half (host half: gz_host.hpp).  Replaces popen("gzip > file") (pbsim.cpp:708-730).
//
// A record stream is cut into UNITS of kGzUnit input bytes; every unit becomes one gzip member:
//   header (10 bytes, FNAME flag) | name: k padding characters + NUL | one dynamic-Huffman DEFLATE block | CRC-32 | ISIZE
// k in 0..3 makes the member's size a multiple of 4, so every member starts 4-byte aligned in the output and a
// member built in shared memory is copied out with whole-word, coalesced stores.
//   k_gz_hist   byte histogram of the stream (the code is built from it on the host)
//   k_gz_size   per unit: exact member size (code lengths are known) and CRC-32 of its text
//   (scan)      member offsets
//   k_gz_encode per unit: one CTA, 256 threads x 128-byte slices; a CTA-wide scan of the slices' bit counts gives
//               every thread the bit position of its codes, which it ORs into the shared-memory image
// Both passes read the text once (4.1 B per emitted base each); the encoder writes about a third of that.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gz_host.hpp"

namespace pb {

constexpr uint32_t kGzUnit = 32768;                        // input bytes per member
constexpr uint32_t kGzThreads = 256;
constexpr uint32_t kGzSlice = kGzUnit / kGzThreads;        // 128 bytes per thread
constexpr uint32_t kGzImgBytes = 24 + kGzHdrWords * 4 + kGzUnit / 8 * kGzMaxBits + 16;  // worst-case member
constexpr uint32_t kGzImgWords = (kGzImgBytes + 3) / 4;

struct GzTables {        // device copy of GzCode + CRC constants
  uint32_t lit[256];
  uint32_t crc[256];
  uint32_t hdr[kGzHdrWords];
  uint32_t x2n[32];      // x^(2^k) mod P
  uint32_t tail[kGzThreads];  // x^(8 * kGzSlice * (kGzThreads-1-t)) mod P: moves slice t's CRC to the end of a full unit
  uint32_t eob, hdr_bits;
  uint8_t len[256];
};

__global__ void k_gz_hist(const uint8_t *__restrict__ in, uint64_t n, uint32_t sample, unsigned long long *hist) {
  __shared__ unsigned int h[256];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  // a sample is enough (every byte value gets a code whatever the counts): 16 bytes per thread out of every
  // `sample` chunks, spread over the whole stream
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16u * sample;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16u * sample; i < n; i += stride) {
    if (i + 16u <= n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + i));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int b = 0; b < 4; ++b) atomicAdd(&h[(w[q] >> (8 * b)) & 0xFFu], 1u);
    } else {
      for (uint64_t j = i; j < n; ++j) atomicAdd(&h[in[j]], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
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

// per-thread: bit count and CRC-32 of its slice
__device__ __forceinline__ void gz_slice_scan(const uint8_t *__restrict__ in, uint64_t pos, uint32_t nbytes,
                                              const uint8_t *len_s, const uint32_t *crc_s, uint32_t *bits,
                                              uint32_t *crc_out) {
  uint32_t nb = 0, crc = 0xFFFFFFFFu;
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
          nb += len_s[c];
          crc = crc_s[(crc ^ c) & 0xFFu] ^ (crc >> 8);
        }
    }
  } else {
    for (uint32_t q = 0; q < nbytes; ++q) {
      const uint32_t c = in[pos + q];
      nb += len_s[c];
      crc = crc_s[(crc ^ c) & 0xFFu] ^ (crc >> 8);
    }
  }
  *bits = nb;
  *crc_out = nbytes ? ~crc : 0u;  // finalised CRC of the slice (0 for an empty slice)
}

// unit_size[u] = member bytes, unit_crc[u] = CRC-32 of the unit's text.  The input is padded to a multiple of 16.
__global__ void __launch_bounds__(kGzThreads) k_gz_size(const uint8_t *__restrict__ in, uint64_t n, const GzTables *T,
                                                        uint32_t bgzf, uint32_t *unit_size, uint32_t *unit_crc) {
  __shared__ uint8_t len_s[256];
  __shared__ uint32_t crc_s[256];
  __shared__ uint32_t red[kGzThreads];
  __shared__ uint32_t wt[8];
  len_s[threadIdx.x] = T->len[threadIdx.x];
  crc_s[threadIdx.x] = T->crc[threadIdx.x];
  __syncthreads();
  const uint64_t u0 = (uint64_t)blockIdx.x * kGzUnit;
  const uint32_t un = (uint32_t)min((uint64_t)kGzUnit, n - u0);
  const uint32_t s0 = threadIdx.x * kGzSlice;
  const uint32_t sn = s0 >= un ? 0u : min(kGzSlice, un - s0);
  uint32_t bits, crc;
  gz_slice_scan(in, u0 + s0, sn, len_s, crc_s, &bits, &crc);
  uint32_t total_bits;
  gz_block_excl_scan(bits, wt, &total_bits);
  // crc(A || B) = crc(A) * x^(8 |B|) + crc(B)  =>  crc(unit) = XOR over slices of crc(slice) * x^(8 * bytes after it):
  // one multiplication per thread (a table constant for full units), then an XOR reduction
  const uint32_t after = un - (s0 + sn);  // bytes of the unit behind this slice
  uint32_t part = 0;
  if (sn != 0u) part = gz_gf_mul_dev(un == kGzUnit ? T->tail[threadIdx.x] : gz_x8n(T->x2n, after), crc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, o);
  if ((threadIdx.x & 31u) == 0u) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t c = 0;
    for (uint32_t q = 0; q < kGzThreads / 32u; ++q) c ^= red[q];
    red[0] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t B, k, hb, total;
    gz_geometry(T->hdr_bits + total_bits + (T->eob >> 16), bgzf, &B, &k, &hb, &total);
    unit_size[blockIdx.x] = total;
    unit_crc[blockIdx.x] = red[0];
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
                                                          uint32_t bgzf, uint8_t *out) {
  extern __shared__ __align__(16) uint32_t gz_smem[];
  uint32_t *img = gz_smem;                       // [kGzImgWords]
  uint32_t *lit_s = gz_smem + kGzImgWords;       // [256]
  __shared__ uint32_t wt[8];
  for (uint32_t i = threadIdx.x; i < kGzImgWords; i += kGzThreads) img[i] = 0;
  lit_s[threadIdx.x] = T->lit[threadIdx.x];
  __syncthreads();
  const uint64_t u0 = (uint64_t)blockIdx.x * kGzUnit;
  const uint32_t un = (uint32_t)min((uint64_t)kGzUnit, n - u0);
  const uint32_t s0 = threadIdx.x * kGzSlice;
  const uint32_t sn = s0 >= un ? 0u : min(kGzSlice, un - s0);
  // pass A over the slice: its bit count
  uint32_t bits = 0;
  if (sn == kGzSlice) {
#pragma unroll 2
    for (uint32_t q = 0; q < kGzSlice; q += 16u) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + u0 + s0 + q));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) bits += lit_s[(w[a] >> (8 * b)) & 0xFFu] >> 16;
    }
  } else {
    for (uint32_t q = 0; q < sn; ++q) bits += lit_s[in[u0 + s0 + q]] >> 16;
  }
  uint32_t total_bits;
  const uint32_t my_bits = gz_block_excl_scan(bits, wt, &total_bits);
  const uint32_t hdr_bits = T->hdr_bits, eob = T->eob;
  uint32_t B, k, hb, total;
  gz_geometry(hdr_bits + total_bits + (eob >> 16), bgzf, &B, &k, &hb, &total);
  const uint32_t body0 = hb * 8u;  // bit position of the DEFLATE stream inside the member
  // header bytes, block header, end of block, trailer
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
    gz_put_bits(img, body0 + hdr_bits + total_bits, eob & 0xFFFFu, eob >> 16);
    const uint32_t tr = (hb + B) * 8u;
    gz_put_bits(img, tr, unit_crc[blockIdx.x], 32);
    gz_put_bits(img, tr + 32u, un, 32);           // ISIZE
  }
  for (uint32_t i = threadIdx.x; i * 32u < hdr_bits; i += kGzThreads) {
    const uint32_t nb = min(32u, hdr_bits - i * 32u);
    gz_put_bits(img, body0 + i * 32u, nb == 32u ? T->hdr[i] : (T->hdr[i] & ((1u << nb) - 1u)), nb);
  }
  // pass B: the slice's codes
  {
    uint32_t pos = body0 + hdr_bits + my_bits;
    uint32_t w = pos >> 5;
    uint32_t accbits = pos & 31u;
    unsigned long long acc = 0ull;
    auto push = [&](uint32_t c) {
      const uint32_t e = lit_s[c];
      acc |= (unsigned long long)(e & 0xFFFFu) << accbits;
      accbits += e >> 16;
      if (accbits >= 32u) {
        atomicOr(&img[w], (uint32_t)acc);
        ++w;
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
    if (accbits) atomicOr(&img[w], (uint32_t)acc);
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
