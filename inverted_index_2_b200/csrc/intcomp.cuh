// intcomp.cuh — device twin of oracle/intcomp_ref.c: the uint32 posting codec behind
// <key>_val (ronanh/intcomp v1.1.0; call sites file/writer.go:49, file/reader.go:100).
// Layout (restated, byte parity UNPINNED — see oracle/intcomp_ref.c for confidence levels):
//   stream  := [binpack-section][varbyte-section]
//   binpack := count(mult. of 128) words(section length incl. 3 header words) first(in[0])
//              then per 128-block: hdr = s1<<31|w1<<24|s2<<23|w2<<16|s3<<15|w3<<8|s4<<7|w4,
//              four groups of 32 deltas at w_i bits, LSB-first, zig-zag iff a delta < 0
//   varbyte := count(1..127) then zigzag(delta) 7 bits/byte, low first, last byte |= 0x80,
//              prev starts at 0, bytes packed little-endian into words, zero padded.
// One list = one independent stream; a 32-group maps onto the 32 lanes of a warp.
#pragma once
#include "common.cuh"

namespace ii2 {
namespace intcomp {

__device__ __forceinline__ uint32_t zigzag(uint32_t cur, uint32_t prev) {
  int32_t d = (int32_t)(cur - prev);
  return ((uint32_t)d << 1) ^ (uint32_t)(d >> 31);
}
__device__ __forceinline__ uint32_t unzigzag(uint32_t z) { return (z >> 1) ^ (0u - (z & 1u)); }
__device__ __forceinline__ int bitlen(uint32_t x) { return 32 - __clz(x); }
__device__ __forceinline__ uint32_t vbyte_len(uint32_t z) {
  return z < (1u << 7) ? 1u : z < (1u << 14) ? 2u : z < (1u << 21) ? 3u : z < (1u << 28) ? 4u : 5u;
}

// Worst-case words for n values.
__host__ __device__ __forceinline__ uint64_t enc_bound(uint64_t n) {
  return 3 + (n / 128) * 129 + 1 + (5 * (n % 128) + 3) / 4 + 1;
}

// ---- size pass: words CompressUint32 would emit for v[0..n).  One warp; uniform result.
// v may live in global or shared memory.
__device__ __forceinline__ uint32_t enc_size_warp(const uint32_t* v, uint32_t n) {
  if (n == 0) return 0;
  const unsigned lane = lane_id();
  const uint32_t nb = n >> 7, r = n & 127u;
  uint32_t words = nb ? 3u : 0u;
  for (uint32_t b = 0; b < nb; b++) {
    words += 1;
#pragma unroll
    for (int g = 0; g < 4; g++) {
      uint32_t idx = b * 128 + g * 32 + lane;
      uint32_t cur = v[idx];
      uint32_t prev = idx ? v[idx - 1] : cur;
      uint32_t m = __reduce_or_sync(0xffffffffu, zigzag(cur, prev));
      words += (m & 1u) ? bitlen(m) : bitlen(m >> 1);
    }
  }
  if (r) {
    uint32_t bytes = 0;
    for (uint32_t i = lane; i < r; i += 32) {
      uint32_t idx = nb * 128 + i;
      uint32_t prev = i ? v[idx - 1] : 0u;
      bytes += vbyte_len(zigzag(v[idx], prev));
    }
    bytes = warp_sum(bytes);
    words += 1 + (bytes + 3) / 4;
  }
  return words;
}

constexpr int kStageWords = 160;  // per-warp shared staging: 32 words (bit packing) / 635 bytes (varbyte)

// ---- emit pass: writes the stream for v[0..n) to dst (global), returns words written
// (uniform).  `stage` = kStageWords of shared memory private to the calling warp.
__device__ __forceinline__ uint32_t enc_emit_warp(const uint32_t* v, uint32_t n, uint32_t* dst,
                                                  uint32_t* stage) {
  if (n == 0) return 0;
  const unsigned lane = lane_id();
  const uint32_t nb = n >> 7, r = n & 127u;
  uint32_t pos = 0;
  if (nb) {
    pos = 3;
    for (uint32_t b = 0; b < nb; b++) {
      uint32_t coded[4];
      int w[4];
      uint32_t hdr = 0;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint32_t idx = b * 128 + g * 32 + lane;
        uint32_t cur = v[idx];
        uint32_t prev = idx ? v[idx - 1] : cur;
        uint32_t z = zigzag(cur, prev);
        uint32_t m = __reduce_or_sync(0xffffffffu, z);
        uint32_t s = m & 1u;
        w[g] = s ? bitlen(m) : bitlen(m >> 1);
        coded[g] = s ? z : (cur - prev);
        hdr |= ((s << 7) | (uint32_t)w[g]) << (24 - 8 * g);
      }
      if (lane == 0) dst[pos] = hdr;
      pos += 1;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const int wd = w[g];
        if (wd == 32) {
          dst[pos + lane] = coded[g];
        } else if (wd > 0) {
          if ((int)lane < wd) stage[lane] = 0;
          __syncwarp();
          uint32_t bit = lane * (uint32_t)wd, sh = bit & 31u;
          atomicOr(&stage[bit >> 5], coded[g] << sh);
          if (sh + (uint32_t)wd > 32u) atomicOr(&stage[(bit >> 5) + 1], coded[g] >> (32u - sh));
          __syncwarp();
          if ((int)lane < wd) dst[pos + lane] = stage[lane];
          __syncwarp();
        }
        pos += (uint32_t)wd;
      }
    }
    if (lane == 0) {
      dst[0] = nb * 128;
      dst[1] = pos;
      dst[2] = v[0];
    }
  }
  if (r) {
    if (lane == 0) dst[pos] = r;
    pos += 1;
    uint8_t* sb = reinterpret_cast<uint8_t*>(stage);
    uint32_t bo = 0;
    for (uint32_t t = 0; t < r; t += 32) {
      uint32_t i = t + lane;
      uint32_t z = 0, len = 0;
      if (i < r) {
        uint32_t idx = nb * 128 + i;
        uint32_t prev = i ? v[idx - 1] : 0u;
        z = zigzag(v[idx], prev);
        len = vbyte_len(z);
      }
      uint32_t inc = warp_inclusive_scan(len);
      uint32_t off = bo + inc - len;
      for (uint32_t k = 0; k < len; k++) {
        uint32_t byte = (z >> (7 * k)) & 0x7Fu;
        if (k + 1 == len) byte |= 0x80u;
        sb[off + k] = (uint8_t)byte;
      }
      bo += __shfl_sync(0xffffffffu, inc, 31);
    }
    uint32_t nwords = (bo + 3) / 4;
    if (lane < nwords * 4 - bo) sb[bo + lane] = 0;  // zero padding of the last word
    __syncwarp();
    for (uint32_t j = lane; j < nwords; j += 32) dst[pos + j] = stage[j];
    __syncwarp();
    pos += nwords;
  }
  return pos;
}

// ---- CTA-cooperative encoder for long lists ------------------------------------------
// words of 128-block `blk` of v (header word + four packed groups); uniform over the warp
__device__ __forceinline__ uint32_t enc_block_size_warp(const uint32_t* v, uint32_t blk) {
  const unsigned lane = lane_id();
  uint32_t words = 1;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const uint32_t idx = blk * 128 + g * 32 + lane;
    const uint32_t cur = v[idx];
    const uint32_t prev = idx ? v[idx - 1] : cur;
    const uint32_t m = __reduce_or_sync(0xffffffffu, zigzag(cur, prev));
    words += (m & 1u) ? bitlen(m) : bitlen(m >> 1);
  }
  return words;
}

// writes 128-block `blk` of v at dst (header + groups); `stage` = 32 words private to the warp
__device__ __forceinline__ void enc_block_emit_warp(const uint32_t* v, uint32_t blk, uint32_t* dst,
                                                    uint32_t* stage) {
  const unsigned lane = lane_id();
  uint32_t coded[4];
  int w[4];
  uint32_t hdr = 0;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const uint32_t idx = blk * 128 + g * 32 + lane;
    const uint32_t cur = v[idx];
    const uint32_t prev = idx ? v[idx - 1] : cur;
    const uint32_t z = zigzag(cur, prev);
    const uint32_t m = __reduce_or_sync(0xffffffffu, z);
    const uint32_t s = m & 1u;
    w[g] = s ? bitlen(m) : bitlen(m >> 1);
    coded[g] = s ? z : (cur - prev);
    hdr |= ((s << 7) | (uint32_t)w[g]) << (24 - 8 * g);
  }
  if (lane == 0) dst[0] = hdr;
  uint32_t pos = 1;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const int wd = w[g];
    if (wd == 32) {
      dst[pos + lane] = coded[g];
    } else if (wd > 0) {
      if ((int)lane < wd) stage[lane] = 0;
      __syncwarp();
      const uint32_t bit = lane * (uint32_t)wd, sh = bit & 31u;
      atomicOr(&stage[bit >> 5], coded[g] << sh);
      if (sh + (uint32_t)wd > 32u) atomicOr(&stage[(bit >> 5) + 1], coded[g] >> (32u - sh));
      __syncwarp();
      if ((int)lane < wd) dst[pos + lane] = stage[lane];
      __syncwarp();
    }
    pos += (uint32_t)wd;
  }
}

// The var-byte section of the `tail` (< 128) values after block nb of v, written at dst by one
// warp; `my_stage` = kStageWords of shared memory private to the warp.
__device__ __forceinline__ void enc_tail_warp(const uint32_t* v, uint32_t nb, uint32_t tail,
                                              uint32_t* dst, uint32_t* my_stage) {
  const unsigned lane = lane_id();
  if (lane == 0) dst[0] = tail;
  uint8_t* sb = reinterpret_cast<uint8_t*>(my_stage);
  uint32_t bo = 0;
  for (uint32_t t = 0; t < tail; t += 32) {
    const uint32_t i = t + lane;
    uint32_t z = 0, len = 0;
    if (i < tail) {
      const uint32_t idx = nb * 128 + i;
      z = zigzag(v[idx], i ? v[idx - 1] : 0u);
      len = vbyte_len(z);
    }
    const uint32_t inc = warp_inclusive_scan(len);
    const uint32_t off = bo + inc - len;
    for (uint32_t k = 0; k < len; k++) {
      uint32_t byte = (z >> (7 * k)) & 0x7Fu;
      if (k + 1 == len) byte |= 0x80u;
      sb[off + k] = (uint8_t)byte;
    }
    bo += __shfl_sync(0xffffffffu, inc, 31);
  }
  const uint32_t nwords = (bo + 3) / 4;
  if (lane < nwords * 4 - bo) sb[bo + lane] = 0;
  __syncwarp();
  for (uint32_t j = lane; j < nwords; j += 32) dst[1 + j] = my_stage[j];
}

// Whole-CTA CompressUint32 of v[0..n) (global memory, n >= 128) to dst.  `table` = n/128 words
// of global scratch (block sizes, then their exclusive prefix); `stage` = kStageWords of shared
// memory per warp; `ws` = block-scan scratch (blockDim/32 + 2).  Every thread of the block
// must call.  Returns the stream length in words (uniform).
__device__ __forceinline__ uint32_t enc_emit_cta(const uint32_t* v, uint32_t n, uint32_t* dst,
                                                 uint32_t* table, uint32_t* stage, uint64_t* ws) {
  const unsigned lane = lane_id(), warp = warp_id(), nwarps = blockDim.x >> 5;
  const uint32_t nb = n >> 7, tail = n & 127u;
  for (uint32_t b = warp; b < nb; b += nwarps) {
    const uint32_t wds = enc_block_size_warp(v, b);
    if (lane == 0) table[b] = wds;
  }
  __syncthreads();
  uint64_t run = 0;
  for (uint32_t base = 0; base < nb; base += blockDim.x) {
    const uint32_t b = base + threadIdx.x;
    const uint64_t x = b < nb ? table[b] : 0u;
    uint64_t tot;
    const uint64_t ex = block_exclusive_scan(x, ws, tot);
    if (b < nb) table[b] = (uint32_t)(run + ex);
    run += tot;
  }
  __syncthreads();
  uint32_t* my_stage = stage + warp * kStageWords;
  for (uint32_t b = warp; b < nb; b += nwarps) enc_block_emit_warp(v, b, dst + 3 + table[b], my_stage);
  uint32_t pos = 3 + (uint32_t)run;
  if (warp == 0) {
    if (lane == 0) {
      dst[0] = nb * 128;
      dst[1] = pos;
      dst[2] = v[0];
    }
  }
  uint32_t tail_words = 0;
  if (tail) {  // var-byte tail: small, one warp; every warp computes the size
    uint32_t bytes = 0;
    for (uint32_t i = lane; i < tail; i += 32) {
      const uint32_t idx = nb * 128 + i;
      bytes += vbyte_len(zigzag(v[idx], i ? v[idx - 1] : 0u));
    }
    bytes = warp_sum(bytes);
    tail_words = 1 + (bytes + 3) / 4;
    if (warp == 0) enc_tail_warp(v, nb, tail, dst + pos, my_stage);
  }
  __syncthreads();
  return pos + tail_words;
}

// ---- single-thread variants for lists below one block (n < 128: varbyte section only) --
__device__ __forceinline__ uint32_t enc_size_thread_small(const uint32_t* v, uint32_t n) {
  if (n == 0) return 0;
  uint32_t bytes = 0, prev = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t cur = v[i];
    bytes += vbyte_len(zigzag(cur, prev));
    prev = cur;
  }
  return 1 + (bytes + 3) / 4;
}

__device__ __forceinline__ uint32_t enc_emit_thread_small(const uint32_t* v, uint32_t n,
                                                          uint32_t* dst) {
  if (n == 0) return 0;
  dst[0] = n;
  uint32_t pos = 1, word = 0, nbytes = 0, prev = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t cur = v[i];
    uint32_t z = zigzag(cur, prev);
    prev = cur;
    for (;;) {
      uint32_t byte = z & 0x7Fu;
      z >>= 7;
      if (z == 0) byte |= 0x80u;
      word |= byte << (8 * nbytes);
      if (++nbytes == 4) {
        dst[pos++] = word;
        word = 0;
        nbytes = 0;
      }
      if (byte & 0x80u) break;
    }
  }
  if (nbytes) dst[pos++] = word;
  return pos;
}

// ---- decode -----------------------------------------------------------------------
// Values a stream decodes to, reading section headers only; -1 if malformed.
__device__ __forceinline__ long long dec_count(const uint32_t* w, uint64_t nwords) {
  uint64_t pos = 0;
  long long total = 0;
  while (pos < nwords) {
    uint32_t c = w[pos];
    if (c == 0) return -1;
    if (c >= 128) {
      if ((c & 127u) || pos + 3 > nwords) return -1;
      uint32_t len = w[pos + 1];
      if (len < 3 || pos + len > nwords) return -1;
      total += c;
      pos += len;
    } else {
      return total + c;  // varbyte section is the last one
    }
  }
  return total;
}

// Sequential varbyte-section decode by one thread. w points at the section header.
// Returns 0 ok, -1 corrupt.
__device__ __forceinline__ int dec_varbyte_thread(const uint32_t* w, uint64_t nwords,
                                                  uint32_t* out) {
  uint32_t c = w[0];
  uint64_t nbytes = (nwords - 1) * 4, bp = 0;
  uint32_t prev = 0, word = 0;
  for (uint32_t i = 0; i < c; i++) {
    uint32_t z = 0;
    int shift = 0;
    for (;;) {
      if (bp >= nbytes || shift > 28) return -1;
      if ((bp & 3) == 0) word = w[1 + (bp >> 2)];
      uint32_t byte = (word >> (8 * (bp & 3))) & 0xFFu;
      bp++;
      z |= (byte & 0x7Fu) << shift;
      shift += 7;
      if (byte & 0x80u) break;
    }
    prev += unzigzag(z);
    out[i] = prev;
  }
  return 0;
}

// One 128-value block by one warp.  p = its header word, end = end of the section, prev = the
// value before the block.  EMIT: the values go to out[0..128) and the last one is returned;
// else only the sum of the block's deltas (mod 2^32) is returned.  *bad is set when a width
// or the section bound is violated.
template <bool EMIT>
__device__ __forceinline__ uint32_t dec_block_warp(const uint32_t* w, uint64_t p, uint64_t end,
                                                   uint32_t prev, uint32_t* out, bool* bad) {
  const unsigned lane = lane_id();
  const uint32_t h = w[p++];
  uint32_t acc = 0;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const uint32_t f = (h >> (24 - 8 * g)) & 0xFFu;
    const uint32_t wd = f & 0x7Fu, s = f >> 7;
    if (wd > 32 || p + wd > end) {
      *bad = true;
      return 0;
    }
    uint32_t val = 0;
    if (wd == 32) {
      val = w[p + lane];
    } else if (wd > 0) {
      const uint32_t bit = lane * wd, wi = bit >> 5, sh = bit & 31u;
      val = w[p + wi] >> sh;
      if (sh + wd > 32u) val |= w[p + wi + 1] << (32u - sh);
      val &= (1u << wd) - 1u;
    }
    const uint32_t d = s ? unzigzag(val) : val;
    if (EMIT) {
      const uint32_t value = prev + warp_inclusive_scan(d);
      out[32 * g + lane] = value;
      prev = __shfl_sync(0xffffffffu, value, 31);
    } else {
      acc += d;
    }
    p += wd;
  }
  if (EMIT) return prev;
  return __reduce_add_sync(0xffffffffu, acc);
}

// Whole-stream decode by one warp (any length).  n = expected values (from dec_count).
// Returns 0 ok, -1 corrupt (uniform).
__device__ __forceinline__ int dec_warp(const uint32_t* w, uint64_t nwords, uint32_t* out,
                                        uint64_t n) {
  const unsigned lane = lane_id();
  uint64_t pos = 0, o = 0;
  while (pos < nwords) {
    uint32_t c = w[pos];
    if (c >= 128) {
      uint32_t len = w[pos + 1];
      uint32_t prev = w[pos + 2];
      uint64_t p = pos + 3, end = pos + len;
      if (o + c > n) return -1;
      for (uint32_t b = 0; b < (c >> 7); b++) {
        if (p >= end) return -1;
        uint32_t h = w[p++];
#pragma unroll
        for (int g = 0; g < 4; g++) {
          uint32_t f = (h >> (24 - 8 * g)) & 0xFFu;
          uint32_t wd = f & 0x7Fu, s = f >> 7;
          if (wd > 32 || p + wd > end) return -1;
          uint32_t val = 0;
          if (wd == 32) {
            val = w[p + lane];
          } else if (wd > 0) {
            uint32_t bit = lane * wd, wi = bit >> 5, sh = bit & 31u;
            val = w[p + wi] >> sh;
            if (sh + wd > 32u) val |= w[p + wi + 1] << (32u - sh);
            val &= (1u << wd) - 1u;
          }
          uint32_t d = s ? unzigzag(val) : val;
          uint32_t inc = warp_inclusive_scan(d);
          uint32_t value = prev + inc;
          out[o + lane] = value;
          prev = __shfl_sync(0xffffffffu, value, 31);
          o += 32;
          p += wd;
        }
      }
      pos = end;
    } else {
      if (c == 0 || o + c > n) return -1;
      int rc = 0;
      if (lane == 0) rc = dec_varbyte_thread(w + pos, nwords - pos, out + o);
      rc = __shfl_sync(0xffffffffu, rc, 0);
      return rc;
    }
  }
  return 0;
}

}  // namespace intcomp
}  // namespace ii2
