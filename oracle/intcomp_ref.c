/*
 * intcomp_ref.c — CPU restatement of ronanh/intcomp v1.1.0 (go.mod:10), the codec
 * behind every byte of <key>_val.  TEST INFRASTRUCTURE (see ii2_oracle.h).
 *
 * Reference call sites: intcomp.CompressUint32(tv.Values, nil) file/writer.go:49,
 * intcomp.UncompressUint32(compressed, nil) file/reader.go:100, and the round
 * trips pinned by file/writer_test.go:13-45 (unsorted {10,500,300}, empty list).
 *
 * PARITY UNPINNED at byte level: the module source is not under /root/reference
 * and cannot be fetched; the reference's tests never assert a codec byte.  This
 * file restates the module's published scheme (README: "blocks of 128x32bit …
 * differential coding, ZigZag if a block holds a negative delta, bit packing
 * into the optimal number of bits; trailing input that cannot fit in a block is
 * encoded using variable length integer encoding") with this concrete layout:
 *
 *   stream   := [binpack-section] [varbyte-section]          (n == 0 -> no words)
 *   binpack  := count(u32, multiple of 128) words(u32, section length incl. these
 *               3 header words) first(u32 = in[0])  block*
 *   block    := hdr(u32) g1 g2 g3 g4 ; hdr = s1<<31|w1<<24|s2<<23|w2<<16|s3<<15|w3<<8|s4<<7|w4
 *               g_i = 32 deltas of w_i bits each, LSB-first, exactly w_i words;
 *               delta vs previous value (first delta of the stream vs in[0] => 0);
 *               s_i = 1 iff some int32 delta in the group is negative, then all 32
 *               are zig-zag coded; w_i = bit length of the OR of the coded deltas.
 *   varbyte  := count(u32, 1..127) bytes… ; each value = zigzag(int32(v - prev)),
 *               prev starts at 0 for the section; 7 bits per byte, low group
 *               first, the LAST byte of a value carries 0x80; bytes packed
 *               little-endian into u32 words, zero padded.
 *
 * Confidence: block header / group packing / zig-zag rule ●●○, 3-word section
 * header ●●○, varbyte header word and prev=0 ●○○.  Everything format-specific
 * lives in this file and in inverted_index_2_b200/csrc/intcomp.cuh (the device
 * twin) so a correction is a two-file change.
 */
#include <string.h>

#include "ii2_oracle.h"

static inline uint32_t zigzag32(int32_t d) { return ((uint32_t)d << 1) ^ (uint32_t)(d >> 31); }
static inline int32_t unzigzag32(uint32_t z) { return (int32_t)(z >> 1) ^ -(int32_t)(z & 1); }
static inline int bitlen32(uint32_t x) { return x ? 32 - __builtin_clz(x) : 0; }

size_t orc_intcomp_bound(size_t n) {
  /* binpack: 3 + per block (1 + 128) ; varbyte: 1 + ceil(5*127/4) */
  return 3 + (n / 128) * 129 + 1 + (5 * (n % 128) + 3) / 4 + 1;
}

/* pack 32 values of w bits, LSB-first, into exactly w words */
static void pack32(const uint32_t* v, int w, uint32_t* out) {
  if (w == 0) return;
  if (w == 32) {
    memcpy(out, v, 32 * sizeof(uint32_t));
    return;
  }
  uint64_t acc = 0;
  int nbits = 0, o = 0;
  for (int i = 0; i < 32; i++) {
    acc |= (uint64_t)v[i] << nbits;
    nbits += w;
    if (nbits >= 32) {
      out[o++] = (uint32_t)acc;
      acc >>= 32;
      nbits -= 32;
    }
  }
}

static void unpack32(const uint32_t* in, int w, uint32_t* v) {
  if (w == 0) {
    memset(v, 0, 32 * sizeof(uint32_t));
    return;
  }
  if (w == 32) {
    memcpy(v, in, 32 * sizeof(uint32_t));
    return;
  }
  uint64_t acc = 0;
  int nbits = 0, o = 0;
  const uint32_t mask = (1u << w) - 1u;
  for (int i = 0; i < 32; i++) {
    if (nbits < w) {
      acc |= (uint64_t)in[o++] << nbits;
      nbits += 32;
    }
    v[i] = (uint32_t)acc & mask;
    acc >>= w;
    nbits -= w;
  }
}

size_t orc_intcomp_encode(const uint32_t* in, size_t n, uint32_t* out) {
  if (n == 0) return 0; /* CompressUint32 of an empty slice appends nothing (writer_test.go:15) */
  size_t pos = 0;
  size_t nb = n / 128;
  if (nb > 0) {
    pos = 3;
    uint32_t prev = in[0];
    for (size_t b = 0; b < nb; b++) {
      uint32_t coded[4][32];
      int w[4], s[4];
      for (int g = 0; g < 4; g++) {
        const uint32_t* src = in + b * 128 + g * 32;
        uint32_t m = 0;
        uint32_t p = prev;
        int32_t d[32];
        for (int i = 0; i < 32; i++) {
          d[i] = (int32_t)(src[i] - p);
          p = src[i];
          m |= zigzag32(d[i]);
        }
        s[g] = (int)(m & 1u);
        w[g] = s[g] ? bitlen32(m) : bitlen32(m >> 1);
        for (int i = 0; i < 32; i++) coded[g][i] = s[g] ? zigzag32(d[i]) : (uint32_t)d[i];
        prev = p;
      }
      out[pos++] = ((uint32_t)s[0] << 31) | ((uint32_t)w[0] << 24) | ((uint32_t)s[1] << 23) |
                   ((uint32_t)w[1] << 16) | ((uint32_t)s[2] << 15) | ((uint32_t)w[2] << 8) |
                   ((uint32_t)s[3] << 7) | (uint32_t)w[3];
      for (int g = 0; g < 4; g++) {
        pack32(coded[g], w[g], out + pos);
        pos += (size_t)w[g];
      }
    }
    out[0] = (uint32_t)(nb * 128);
    out[1] = (uint32_t)pos;
    out[2] = in[0];
  }
  size_t r = n - nb * 128;
  if (r > 0) {
    out[pos++] = (uint32_t)r;
    uint32_t prev = 0;
    uint32_t word = 0;
    int nbytes = 0;
    for (size_t i = nb * 128; i < n; i++) {
      uint32_t z = zigzag32((int32_t)(in[i] - prev));
      prev = in[i];
      for (;;) {
        uint32_t byte = z & 0x7Fu;
        z >>= 7;
        if (z == 0) byte |= 0x80u;
        word |= byte << (8 * nbytes);
        if (++nbytes == 4) {
          out[pos++] = word;
          word = 0;
          nbytes = 0;
        }
        if (byte & 0x80u) break;
      }
    }
    if (nbytes) out[pos++] = word;
  }
  return pos;
}

size_t orc_intcomp_count(const uint32_t* words, size_t nwords) {
  size_t pos = 0, total = 0;
  while (pos < nwords) {
    uint32_t c = words[pos];
    if (c == 0) return (size_t)-1;
    if (c >= 128) {
      if ((c & 127u) || pos + 3 > nwords) return (size_t)-1;
      uint32_t len = words[pos + 1];
      if (len < 3 || pos + len > nwords) return (size_t)-1;
      total += c;
      pos += len;
    } else {
      /* varbyte section is always the last one produced by CompressUint32 */
      total += c;
      return total;
    }
  }
  return total;
}

size_t orc_intcomp_decode(const uint32_t* words, size_t nwords, uint32_t* out, size_t cap) {
  size_t pos = 0, o = 0;
  while (pos < nwords) {
    uint32_t c = words[pos];
    if (c == 0) return (size_t)-1;
    if (c >= 128) {
      if ((c & 127u) || pos + 3 > nwords) return (size_t)-1;
      uint32_t len = words[pos + 1];
      if (len < 3 || pos + len > nwords || o + c > cap) return (size_t)-1;
      uint32_t prev = words[pos + 2];
      size_t p = pos + 3, end = pos + len;
      for (uint32_t b = 0; b < c / 128; b++) {
        if (p >= end) return (size_t)-1;
        uint32_t h = words[p++];
        int s[4] = {(int)(h >> 31) & 1, (int)(h >> 23) & 1, (int)(h >> 15) & 1, (int)(h >> 7) & 1};
        int w[4] = {(int)(h >> 24) & 0x7F, (int)(h >> 16) & 0x7F, (int)(h >> 8) & 0x7F,
                    (int)h & 0x7F};
        for (int g = 0; g < 4; g++) {
          if (w[g] > 32 || p + (size_t)w[g] > end) return (size_t)-1;
          uint32_t v[32];
          unpack32(words + p, w[g], v);
          p += (size_t)w[g];
          for (int i = 0; i < 32; i++) {
            int32_t d = s[g] ? unzigzag32(v[i]) : (int32_t)v[i];
            prev += (uint32_t)d;
            out[o++] = prev;
          }
        }
      }
      pos = end;
    } else {
      if (o + c > cap) return (size_t)-1;
      const uint8_t* bytes = (const uint8_t*)(words + pos + 1);
      size_t nbytes = (nwords - pos - 1) * 4, bp = 0;
      uint32_t prev = 0;
      for (uint32_t i = 0; i < c; i++) {
        uint32_t z = 0;
        int shift = 0;
        for (;;) {
          if (bp >= nbytes || shift > 28) return (size_t)-1;
          uint32_t byte = bytes[bp++];
          z |= (byte & 0x7Fu) << shift;
          shift += 7;
          if (byte & 0x80u) break;
        }
        prev += (uint32_t)unzigzag32(z);
        out[o++] = prev;
      }
      return o; /* last section */
    }
  }
  return o;
}

/* ---- batched helpers (one CompressUint32 / UncompressUint32 per list, the way
 *      Writer.Append / Reader.Next call them, file/writer.go:49, file/reader.go:100) */
uint64_t orc_intcomp_encode_batch(const uint32_t* in, const uint64_t* off, uint64_t nlists,
                                  uint32_t* out, uint64_t* word_off) {
  uint64_t pos = 0;
  word_off[0] = 0;
  for (uint64_t i = 0; i < nlists; i++) {
    pos += orc_intcomp_encode(in + off[i], (size_t)(off[i + 1] - off[i]), out + pos);
    word_off[i + 1] = pos;
  }
  return pos;
}

/* counts[i+1] = running total of decoded values; returns total or (uint64_t)-1 */
uint64_t orc_intcomp_count_batch(const uint32_t* words, const uint64_t* word_off, uint64_t nlists,
                                 uint64_t* out_off) {
  uint64_t total = 0;
  out_off[0] = 0;
  for (uint64_t i = 0; i < nlists; i++) {
    size_t c = orc_intcomp_count(words + word_off[i], (size_t)(word_off[i + 1] - word_off[i]));
    if (c == (size_t)-1) return (uint64_t)-1;
    total += c;
    out_off[i + 1] = total;
  }
  return total;
}

int orc_intcomp_decode_batch(const uint32_t* words, const uint64_t* word_off, uint64_t nlists,
                             const uint64_t* out_off, uint32_t* out) {
  for (uint64_t i = 0; i < nlists; i++) {
    size_t n = (size_t)(out_off[i + 1] - out_off[i]);
    size_t got = orc_intcomp_decode(words + word_off[i], (size_t)(word_off[i + 1] - word_off[i]),
                                    out + out_off[i], n);
    if (got != n) return II2_ERR_CORRUPT;
  }
  return II2_OK;
}
