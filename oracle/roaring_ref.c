/*
 * roaring_ref.c — CPU restatement of file/bitmask.go (Bitmask[uint32]) and of the
 * parts of RoaringBitmap/roaring v1.9.4 (go.mod:6) it calls.  TEST INFRASTRUCTURE.
 *
 *   Bitmask.Put      file/bitmask.go:53-59   BitmapOf(); Add(indexOf(v)) per value; ToBytes()
 *   Bitmask.indexOf  file/bitmask.go:64-71   slices.Index, append on miss
 *   Bitmask.Get      file/bitmask.go:30-49   ReadFrom(one bitmap), ascending Iterator,
 *                                            values[idx], "bitmask is out of bound"
 *   known answers    file/bitmask_test.go:34-52
 *
 * roaring behaviour restated (module not vendored; bytes PARITY UNPINNED, format is
 * the public RoaringFormatSpec ●●●; container-type rules ●●○):
 *   - Add(): array container while cardinality <= 4096; the 4097th distinct value
 *     converts it to a bitmap container; a bitmap container never shrinks back; a
 *     bitmap container that becomes full (65536) is replaced by the run container
 *     [0,65535].  RunOptimize is never called by Bitmask.
 *   - ToBytes(): no run container -> cookie 12346, container count; else cookie
 *     12347 | (count-1)<<16 followed by ceil(count/8) is-run flag bytes.  Then
 *     (key u16, cardinality-1 u16) per container; then u32 byte offsets per container
 *     (always without runs; with runs only if count >= 4); then payloads: array =
 *     sorted u16s, bitmap = 1024 u64, run = n_runs u16 + (start u16, length-1 u16)*.
 */
#include <stdlib.h>
#include <string.h>

#include "ii2_oracle.h"

enum { C_ARRAY = 0, C_BITMAP = 1, C_RUN = 2 };

typedef struct {
  uint16_t key;
  int type;
  uint32_t card;
  uint16_t* arr; /* C_ARRAY: sorted, cap 4096 */
  uint64_t* bits; /* C_BITMAP: 1024 words */
} rcontainer;

typedef struct {
  rcontainer* c;
  size_t n, cap;
} rbitmap;

static rcontainer* rb_get_container(rbitmap* rb, uint16_t key) {
  /* containers kept sorted by key, as roaring's roaringArray does */
  size_t lo = 0, hi = rb->n;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (rb->c[mid].key < key)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo < rb->n && rb->c[lo].key == key) return &rb->c[lo];
  if (rb->n == rb->cap) {
    rb->cap = rb->cap ? rb->cap * 2 : 4;
    rb->c = (rcontainer*)realloc(rb->c, rb->cap * sizeof(rcontainer));
  }
  memmove(&rb->c[lo + 1], &rb->c[lo], (rb->n - lo) * sizeof(rcontainer));
  rb->n++;
  rcontainer* c = &rb->c[lo];
  c->key = key;
  c->type = C_ARRAY;
  c->card = 0;
  c->arr = (uint16_t*)malloc(4096 * sizeof(uint16_t));
  c->bits = NULL;
  return c;
}

static void rb_add(rbitmap* rb, uint32_t x) {
  rcontainer* c = rb_get_container(rb, (uint16_t)(x >> 16));
  uint16_t lowbits = (uint16_t)x;
  if (c->type == C_RUN) return; /* already full */
  if (c->type == C_BITMAP) {
    uint64_t bit = 1ull << (lowbits & 63);
    if (!(c->bits[lowbits >> 6] & bit)) {
      c->bits[lowbits >> 6] |= bit;
      if (++c->card == 65536) { /* full bitmap container -> run [0,65535] */
        free(c->bits);
        c->bits = NULL;
        c->type = C_RUN;
      }
    }
    return;
  }
  /* array container: binary search, insert keeping order */
  uint32_t lo = 0, hi = c->card;
  while (lo < hi) {
    uint32_t mid = (lo + hi) / 2;
    if (c->arr[mid] < lowbits)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo < c->card && c->arr[lo] == lowbits) return;
  if (c->card >= 4096) { /* arrayDefaultMaxSize reached: convert, then add */
    c->bits = (uint64_t*)calloc(1024, sizeof(uint64_t));
    for (uint32_t i = 0; i < c->card; i++) c->bits[c->arr[i] >> 6] |= 1ull << (c->arr[i] & 63);
    free(c->arr);
    c->arr = NULL;
    c->type = C_BITMAP;
    c->bits[lowbits >> 6] |= 1ull << (lowbits & 63);
    c->card++;
    return;
  }
  memmove(&c->arr[lo + 1], &c->arr[lo], (c->card - lo) * sizeof(uint16_t));
  c->arr[lo] = lowbits;
  c->card++;
}

static void rb_free(rbitmap* rb) {
  for (size_t i = 0; i < rb->n; i++) {
    free(rb->c[i].arr);
    free(rb->c[i].bits);
  }
  free(rb->c);
}

static inline void put16(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
}
static inline void put32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
  p[2] = (uint8_t)(v >> 16);
  p[3] = (uint8_t)(v >> 24);
}
static inline uint32_t get16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static inline uint32_t get32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

static uint8_t* rb_to_bytes(const rbitmap* rb, uint64_t* nbytes) {
  size_t n = rb->n;
  int has_run = 0;
  for (size_t i = 0; i < n; i++) has_run |= (rb->c[i].type == C_RUN);
  size_t hdr = has_run ? 4 + (n + 7) / 8 : 8;
  size_t desc = 4 * n;
  int has_off = !has_run || n >= 4;
  size_t offs = has_off ? 4 * n : 0;
  size_t total = hdr + desc + offs;
  for (size_t i = 0; i < n; i++) {
    const rcontainer* c = &rb->c[i];
    total += c->type == C_ARRAY ? 2 * (size_t)c->card : c->type == C_BITMAP ? 8192 : 2 + 4;
  }
  uint8_t* buf = (uint8_t*)calloc(total ? total : 1, 1);
  size_t p = 0;
  if (has_run) {
    put32(buf, 12347u | ((uint32_t)(n - 1) << 16));
    p = 4;
    for (size_t i = 0; i < n; i++)
      if (rb->c[i].type == C_RUN) buf[p + i / 8] |= (uint8_t)(1u << (i % 8));
    p += (n + 7) / 8;
  } else {
    put32(buf, 12346u);
    put32(buf + 4, (uint32_t)n);
    p = 8;
  }
  for (size_t i = 0; i < n; i++) {
    put16(buf + p, rb->c[i].key);
    put16(buf + p + 2, rb->c[i].card - 1);
    p += 4;
  }
  size_t data = p + offs;
  if (has_off) {
    size_t o = data;
    for (size_t i = 0; i < n; i++) {
      put32(buf + p, (uint32_t)o);
      p += 4;
      const rcontainer* c = &rb->c[i];
      o += c->type == C_ARRAY ? 2 * (size_t)c->card : c->type == C_BITMAP ? 8192 : 6;
    }
  }
  p = data;
  for (size_t i = 0; i < n; i++) {
    const rcontainer* c = &rb->c[i];
    if (c->type == C_ARRAY) {
      for (uint32_t k = 0; k < c->card; k++) put16(buf + p + 2 * k, c->arr[k]);
      p += 2 * (size_t)c->card;
    } else if (c->type == C_BITMAP) {
      for (int k = 0; k < 1024; k++) {
        put32(buf + p + 8 * k, (uint32_t)c->bits[k]);
        put32(buf + p + 8 * k + 4, (uint32_t)(c->bits[k] >> 32));
      }
      p += 8192;
    } else {
      put16(buf + p, 1);
      put16(buf + p + 2, 0);
      put16(buf + p + 4, 65535);
      p += 6;
    }
  }
  *nbytes = total;
  return buf;
}

/* ---- Bitmask ------------------------------------------------------------- */
struct orc_bitmask {
  uint32_t* values;
  uint64_t n, cap;
  /* optional value -> first index map (open addressing), only for fast mode */
  uint64_t* map; /* (value<<32 | index+1), 0 = empty */
  uint64_t map_cap, map_n;
};

static void bm_push(orc_bitmask* bm, uint32_t v) {
  if (bm->n == bm->cap) {
    bm->cap = bm->cap ? bm->cap * 2 : 16;
    bm->values = (uint32_t*)realloc(bm->values, bm->cap * sizeof(uint32_t));
  }
  bm->values[bm->n++] = v;
}

static inline uint64_t hash32(uint32_t x) {
  uint64_t h = (uint64_t)x * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}

static void map_insert_raw(uint64_t* map, uint64_t cap, uint32_t v, uint64_t idx) {
  uint64_t h = hash32(v) & (cap - 1);
  for (;;) {
    if (map[h] == 0) {
      map[h] = ((uint64_t)v << 32) | (idx + 1);
      return;
    }
    if ((uint32_t)(map[h] >> 32) == v) return; /* keep FIRST index */
    h = (h + 1) & (cap - 1);
  }
}

static void map_build(orc_bitmask* bm) {
  uint64_t cap = 64;
  while (cap < 2 * (bm->n + 1)) cap <<= 1;
  bm->map = (uint64_t*)calloc(cap, sizeof(uint64_t));
  bm->map_cap = cap;
  bm->map_n = bm->n;
  for (uint64_t i = 0; i < bm->n; i++) {
    /* index+1 stored in low 32 bits: dictionary must stay below 2^32-1 entries */
    map_insert_raw(bm->map, cap, bm->values[i], i);
  }
}

static int64_t map_find(const orc_bitmask* bm, uint32_t v) {
  uint64_t h = hash32(v) & (bm->map_cap - 1);
  for (;;) {
    uint64_t e = bm->map[h];
    if (e == 0) return -1;
    if ((uint32_t)(e >> 32) == v) return (int64_t)(e & 0xFFFFFFFFu) - 1;
    h = (h + 1) & (bm->map_cap - 1);
  }
}

orc_bitmask* orc_bitmask_new(const uint32_t* init, uint64_t n) {
  orc_bitmask* bm = (orc_bitmask*)calloc(1, sizeof(orc_bitmask));
  for (uint64_t i = 0; i < n; i++) bm_push(bm, init[i]);
  return bm;
}

void orc_bitmask_free(orc_bitmask* bm) {
  if (!bm) return;
  free(bm->values);
  free(bm->map);
  free(bm);
}

uint64_t orc_bitmask_len(const orc_bitmask* bm) { return bm->n; }
const uint32_t* orc_bitmask_values(const orc_bitmask* bm) { return bm->values; }

/* file/bitmask.go:64-71 */
static uint32_t bm_index_of(orc_bitmask* bm, uint32_t v, int fast) {
  if (fast) {
    if (!bm->map || 2 * (bm->n + 1) > bm->map_cap) {
      free(bm->map);
      map_build(bm);
    }
    int64_t pos = map_find(bm, v);
    if (pos < 0) {
      bm_push(bm, v);
      pos = (int64_t)bm->n - 1;
      map_insert_raw(bm->map, bm->map_cap, v, (uint64_t)pos);
    }
    return (uint32_t)pos;
  }
  for (uint64_t i = 0; i < bm->n; i++) /* slices.Index */
    if (bm->values[i] == v) return (uint32_t)i;
  bm_push(bm, v);
  if (bm->map) { /* keep an existing fast map coherent */
    free(bm->map);
    bm->map = NULL;
  }
  return (uint32_t)(bm->n - 1);
}

/* file/bitmask.go:53-59 */
int orc_bitmask_put(orc_bitmask* bm, const uint32_t* vals, uint64_t n, int fast, uint8_t** bytes,
                    uint64_t* nbytes) {
  rbitmap rb = {0};
  for (uint64_t i = 0; i < n; i++) rb_add(&rb, bm_index_of(bm, vals[i], fast));
  *bytes = rb_to_bytes(&rb, nbytes);
  rb_free(&rb);
  return II2_OK;
}

/* file/bitmask.go:30-49; parses exactly one bitmap from the front of enc */
int orc_bitmask_get(const orc_bitmask* bm, const uint8_t* enc, uint64_t nenc, uint32_t** vals,
                    uint64_t* n) {
  *vals = NULL;
  *n = 0;
  if (nenc < 4) return II2_ERR_CORRUPT;
  uint32_t cookie = get32(enc);
  size_t p, nc;
  const uint8_t* runflags = NULL;
  int has_run = 0;
  if ((cookie & 0xFFFF) == 12347) {
    has_run = 1;
    nc = (cookie >> 16) + 1;
    runflags = enc + 4;
    p = 4 + (nc + 7) / 8;
  } else if (cookie == 12346) {
    if (nenc < 8) return II2_ERR_CORRUPT;
    nc = get32(enc + 4);
    p = 8;
  } else {
    return II2_ERR_CORRUPT;
  }
  if (nc > 65536 || p + 4 * nc > nenc) return II2_ERR_CORRUPT;
  const uint8_t* desc = enc + p;
  p += 4 * nc;
  if (!has_run || nc >= 4) p += 4 * nc; /* offset header: skipped, payloads are sequential */
  if (p > nenc) return II2_ERR_CORRUPT;
  uint64_t total = 0;
  for (size_t i = 0; i < nc; i++) total += get16(desc + 4 * i + 2) + 1;
  uint32_t* out = (uint32_t*)malloc((total ? total : 1) * sizeof(uint32_t));
  uint64_t o = 0;
  int rc = II2_OK;
  for (size_t i = 0; i < nc && rc == II2_OK; i++) {
    uint32_t key = get16(desc + 4 * i), card = get16(desc + 4 * i + 2) + 1;
    int is_run = has_run && (runflags[i / 8] >> (i % 8) & 1);
#define EMIT(idx_)                                              \
  do {                                                          \
    uint32_t idx__ = (idx_);                                    \
    if (idx__ >= bm->n) { /* file/bitmask.go:41-44 */           \
      rc = II2_ERR_BITMASK_OOB;                                 \
    } else {                                                    \
      out[o++] = bm->values[idx__];                             \
    }                                                           \
  } while (0)
    if (is_run) {
      if (p + 2 > nenc) { rc = II2_ERR_CORRUPT; break; }
      uint32_t nr = get16(enc + p);
      p += 2;
      if (p + 4 * (size_t)nr > nenc) { rc = II2_ERR_CORRUPT; break; }
      for (uint32_t r = 0; r < nr && rc == II2_OK; r++) {
        uint32_t start = get16(enc + p + 4 * r), len = get16(enc + p + 4 * r + 2);
        for (uint32_t k = 0; k <= len && rc == II2_OK; k++) {
          if (o >= total) { rc = II2_ERR_CORRUPT; break; }
          EMIT((key << 16) | (start + k));
        }
      }
      p += 4 * (size_t)nr;
    } else if (card > 4096) {
      if (p + 8192 > nenc) { rc = II2_ERR_CORRUPT; break; }
      for (uint32_t w = 0; w < 2048 && rc == II2_OK; w++) {
        uint32_t bits = get32(enc + p + 4 * w);
        while (bits && rc == II2_OK) {
          uint32_t b = (uint32_t)__builtin_ctz(bits);
          bits &= bits - 1;
          if (o >= total) { rc = II2_ERR_CORRUPT; break; }
          EMIT((key << 16) | (w * 32 + b));
        }
      }
      p += 8192;
    } else {
      if (p + 2 * (size_t)card > nenc) { rc = II2_ERR_CORRUPT; break; }
      for (uint32_t k = 0; k < card && rc == II2_OK; k++) EMIT((key << 16) | get16(enc + p + 2 * k));
      p += 2 * (size_t)card;
    }
#undef EMIT
  }
  if (rc != II2_OK) {
    free(out);
    return rc;
  }
  *vals = out;
  *n = o;
  return II2_OK;
}
