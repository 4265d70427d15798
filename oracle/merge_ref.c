/*
 * merge_ref.c — CPU restatement of the reference's compaction and range-read path.
 * TEST INFRASTRUCTURE (see ii2_oracle.h).  Deliberately keeps the reference's
 * algorithmic shape — pull-style readers with one-term look-ahead, a selection
 * structure over reader heads, PAIRWISE append+sort+compact unions, one binary
 * search per value in the removed list, one codec call per term — because it is
 * also the CPU baseline bench.py times ("C restatement of the Go path").
 *
 *   reader            file/reader.go:33-103 (Next), :136-199 (NewReader)
 *   merging iterator  shard.go:253-278 (makeIterator) -> go-iterators MergingIterator
 *                     with file.CompareTermValues / file.MergeTermValues
 *   union             file/types.go:14-22
 *   compare           file/types.go:24-26 (bytes.Compare)
 *   merge loop        shard.go:158-212 (min/max pre-filter :176-179, filter :181-190,
 *                     empty drop :192-194, lazy writer :197-205, termsCount :211)
 *   writer            file/writer.go:32-59 (FST value = running valuesOffset, _val =
 *                     concatenated intcomp words, little-endian)
 *   shardKey          shard.go:362-378
 *
 * Excluded quirk (survey Q1): Reader.Next mis-decodes compressed runs > 16 KiB on a
 * fresh reader (file/reader.go:84-98); parity is defined where the reference is
 * well-defined, so runs of any length decode correctly here.
 */
#include <stdlib.h>
#include <string.h>

#include "ii2_oracle.h"

/* ---- small helpers --------------------------------------------------------- */
typedef struct {
  uint8_t* p;
  size_t n, cap;
} bytebuf;

static int bb_reserve(bytebuf* b, size_t extra) {
  if (b->n + extra <= b->cap) return 0;
  size_t cap = b->cap ? b->cap : 256;
  while (cap < b->n + extra) cap *= 2;
  uint8_t* q = (uint8_t*)realloc(b->p, cap);
  if (!q) return -1;
  b->p = q;
  b->cap = cap;
  return 0;
}
static int bb_append(bytebuf* b, const void* src, size_t n) {
  if (bb_reserve(b, n)) return -1;
  if (n) memcpy(b->p + b->n, src, n);
  b->n += n;
  return 0;
}

/* bytes.Compare */
static int bytes_compare(const uint8_t* a, size_t na, const uint8_t* b, size_t nb) {
  size_t m = na < nb ? na : nb;
  int c = m ? memcmp(a, b, m) : 0;
  if (c) return c < 0 ? -1 : 1;
  return na < nb ? -1 : na > nb ? 1 : 0;
}

/* slices.Sort on uint32 (pattern-defeating quicksort in Go; a plain introsort-style
 * comparison sort here — same asymptotics, same result) */
static void sort_u32(uint32_t* a, size_t n) {
  while (n > 24) {
    uint32_t x = a[0], y = a[n / 2], z = a[n - 1];
    uint32_t piv = x < y ? (y < z ? y : (x < z ? z : x)) : (x < z ? x : (y < z ? z : y));
    size_t i = 0, j = n - 1;
    for (;;) {
      while (a[i] < piv) i++;
      while (a[j] > piv) j--;
      if (i >= j) break;
      uint32_t t = a[i];
      a[i] = a[j];
      a[j] = t;
      i++;
      j--;
    }
    /* recurse on the smaller part */
    size_t left = j + 1, right = n - left;
    if (left < right) {
      sort_u32(a, left);
      a += left;
      n = right;
    } else {
      sort_u32(a + left, right);
      n = left;
    }
  }
  for (size_t i = 1; i < n; i++) {
    uint32_t v = a[i];
    size_t j = i;
    while (j > 0 && a[j - 1] > v) {
      a[j] = a[j - 1];
      j--;
    }
    a[j] = v;
  }
}

/* ---- TermValues (file/types.go:9-12) -------------------------------------- */
typedef struct {
  const uint8_t* term;
  uint32_t term_len;
  uint32_t* values; /* owned */
  size_t nvalues;
} term_values;

/* file.MergeTermValues, file/types.go:14-22: append, sort, compact, fresh copy */
static int merge_term_values(term_values* a, term_values* b) {
  size_t n = a->nvalues + b->nvalues;
  uint32_t* u = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  if (!u) return -1;
  if (a->nvalues) memcpy(u, a->values, a->nvalues * sizeof(uint32_t));
  if (b->nvalues) memcpy(u + a->nvalues, b->values, b->nvalues * sizeof(uint32_t));
  sort_u32(u, n);
  size_t k = 0;
  for (size_t i = 0; i < n; i++)
    if (i == 0 || u[i] != u[i - 1]) u[k++] = u[i];
  uint32_t* fresh = (uint32_t*)malloc((k ? k : 1) * sizeof(uint32_t));
  if (!fresh) {
    free(u);
    return -1;
  }
  if (k) memcpy(fresh, u, k * sizeof(uint32_t));
  free(u);
  free(a->values);
  free(b->values);
  b->values = NULL;
  a->values = fresh;
  a->nvalues = k;
  return 0;
}

/* ---- Reader (file/reader.go) ----------------------------------------------- */
typedef struct {
  const ii2_seg_view* seg;
  uint64_t cur; /* index of prevTerm */
  int done;     /* prevFstError == ErrIteratorDone */
  const uint8_t* max;
  size_t maxlen;
  int has_max;
  /* current head, filled by reader_next */
  term_values head;
  int has_head;
} reader;

static inline const uint8_t* seg_term(const ii2_seg_view* s, uint64_t i, uint32_t* len) {
  *len = s->term_off[i + 1] - s->term_off[i];
  return s->term_bytes + s->term_off[i];
}

/* NewReader, file/reader.go:136-199.  Returns 1 if the segment has no term in range
 * (vellum.ErrIteratorDone, skipped at shard.go:257-261). */
static int reader_open(reader* r, const ii2_seg_view* seg, const uint8_t* min, size_t minlen,
                       int has_min, const uint8_t* max, size_t maxlen, int has_max) {
  memset(r, 0, sizeof(*r));
  r->seg = seg;
  uint64_t lo = 0, hi = seg->n_terms;
  if (has_min) { /* fst.Iterator(min, nil): first key >= min */
    while (lo < hi) {
      uint64_t mid = lo + (hi - lo) / 2;
      uint32_t len;
      const uint8_t* t = seg_term(seg, mid, &len);
      if (bytes_compare(t, len, min, minlen) < 0)
        lo = mid + 1;
      else
        hi = mid;
    }
  }
  if (lo >= seg->n_terms) return 1;
  if (has_max) {
    uint32_t len;
    const uint8_t* t = seg_term(seg, lo, &len);
    if (bytes_compare(t, len, max, maxlen) > 0) return 1; /* :151-155 */
  }
  r->cur = lo;
  r->max = max;
  r->maxlen = maxlen;
  r->has_max = has_max;
  return 0;
}

/* Reader.Next, file/reader.go:33-103.  0 = produced head, 1 = EmptyIterator, <0 error */
static int reader_next(reader* r) {
  r->has_head = 0;
  if (r->done) return 1;
  const ii2_seg_view* s = r->seg;
  uint64_t i = r->cur;
  uint64_t run_size = 0;
  if (i + 1 < s->n_terms) { /* peek succeeded */
    if (s->mode == II2_SEG_VAL) run_size = s->val_off[i + 1] - s->val_off[i]; /* :52 */
    if (r->has_max) {
      uint32_t len;
      const uint8_t* t = seg_term(s, i + 1, &len);
      if (bytes_compare(t, len, r->max, r->maxlen) > 0) r->done = 1; /* :54-58 */
    }
    r->cur = i + 1;
  } else { /* peek hit ErrIteratorDone: run takes the rest of the file, :64 */
    if (s->mode == II2_SEG_VAL) run_size = s->val_size - s->val_off[i];
    r->done = 1;
  }
  term_values* tv = &r->head;
  tv->term = seg_term(s, i, &tv->term_len);
  tv->values = NULL;
  tv->nvalues = 0;
  if (s->mode == II2_SEG_DIRECT) { /* :73-77 */
    tv->values = (uint32_t*)malloc(sizeof(uint32_t));
    if (!tv->values) return II2_ERR_NOMEM;
    tv->values[0] = (uint32_t)s->val_off[i];
    tv->nvalues = 1;
  } else if (s->mode == II2_SEG_DECODED) {
    size_t n = (size_t)(s->post_off[i + 1] - s->post_off[i]);
    tv->values = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    if (!tv->values) return II2_ERR_NOMEM;
    if (n) memcpy(tv->values, s->post + s->post_off[i], n * sizeof(uint32_t));
    tv->nvalues = n;
  } else if (s->mode == II2_SEG_VAL) {
    /* compressed := make([]uint32, runSize/4); binary.Read LE; UncompressUint32, :79-100 */
    size_t nwords = (size_t)(run_size / 4);
    uint32_t* words = (uint32_t*)malloc((nwords ? nwords : 1) * sizeof(uint32_t));
    if (!words) return II2_ERR_NOMEM;
    if (nwords) memcpy(words, s->val_bytes + s->val_off[i], nwords * 4);
    size_t n = orc_intcomp_count(words, nwords);
    if (n == (size_t)-1) {
      free(words);
      return II2_ERR_CORRUPT;
    }
    tv->values = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    if (!tv->values) {
      free(words);
      return II2_ERR_NOMEM;
    }
    size_t got = orc_intcomp_decode(words, nwords, tv->values, n);
    free(words);
    if (got != n) return II2_ERR_CORRUPT;
    tv->nvalues = n;
  } else {
    return II2_ERR_INVALID;
  }
  r->has_head = 1;
  return 0;
}

/* ---- MergingIterator (go-iterators, built at shard.go:267) ------------------- */
typedef struct {
  reader* readers;
  int nreaders;
  int* heap; /* reader indexes ordered by head term (ties: reader index) */
  int nheap;
} merging_iter;

static int head_less(const merging_iter* it, int a, int b) {
  const term_values* x = &it->readers[a].head;
  const term_values* y = &it->readers[b].head;
  int c = bytes_compare(x->term, x->term_len, y->term, y->term_len);
  return c < 0 || (c == 0 && a < b);
}
static void heap_sift_down(merging_iter* it, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < it->nheap && head_less(it, it->heap[l], it->heap[m])) m = l;
    if (r < it->nheap && head_less(it, it->heap[r], it->heap[m])) m = r;
    if (m == i) return;
    int t = it->heap[i];
    it->heap[i] = it->heap[m];
    it->heap[m] = t;
    i = m;
  }
}
static void heap_push(merging_iter* it, int ridx) {
  int i = it->nheap++;
  it->heap[i] = ridx;
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!head_less(it, it->heap[i], it->heap[p])) break;
    int t = it->heap[i];
    it->heap[i] = it->heap[p];
    it->heap[p] = t;
    i = p;
  }
}
static int heap_pop(merging_iter* it) {
  int top = it->heap[0];
  it->heap[0] = it->heap[--it->nheap];
  if (it->nheap) heap_sift_down(it, 0);
  return top;
}

/* makeIterator, shard.go:253-278 */
static int merging_open(merging_iter* it, const ii2_seg_view* segs, int nseg, const uint8_t* min,
                        size_t minlen, int has_min, const uint8_t* max, size_t maxlen,
                        int has_max) {
  memset(it, 0, sizeof(*it));
  it->readers = (reader*)calloc((size_t)(nseg ? nseg : 1), sizeof(reader));
  it->heap = (int*)calloc((size_t)(nseg ? nseg : 1), sizeof(int));
  if (!it->readers || !it->heap) return II2_ERR_NOMEM;
  for (int s = 0; s < nseg; s++) {
    reader* r = &it->readers[it->nreaders];
    if (reader_open(r, &segs[s], min, minlen, has_min, max, maxlen, has_max)) continue;
    int rc = reader_next(r);
    if (rc < 0) return rc;
    if (rc == 0) heap_push(it, it->nreaders);
    it->nreaders++;
  }
  return 0;
}

/* 0 = produced *out (values owned by caller), 1 = EmptyIterator, <0 error */
static int merging_next(merging_iter* it, term_values* out) {
  if (it->nheap == 0) return 1;
  int r0 = heap_pop(it);
  *out = it->readers[r0].head;
  it->readers[r0].has_head = 0;
  /* fold every other head with an equal term, pairwise (file/types.go:14-22) */
  int advanced_cap = 8, nadv = 0;
  int* adv = (int*)malloc((size_t)advanced_cap * sizeof(int));
  if (!adv) return II2_ERR_NOMEM;
  adv[nadv++] = r0;
  while (it->nheap) {
    int top = it->heap[0];
    term_values* h = &it->readers[top].head;
    if (bytes_compare(h->term, h->term_len, out->term, out->term_len) != 0) break;
    heap_pop(it);
    if (merge_term_values(out, h)) {
      free(adv);
      return II2_ERR_NOMEM;
    }
    it->readers[top].has_head = 0;
    if (nadv == advanced_cap) {
      advanced_cap *= 2;
      adv = (int*)realloc(adv, (size_t)advanced_cap * sizeof(int));
    }
    adv[nadv++] = top;
  }
  for (int k = 0; k < nadv; k++) {
    int rc = reader_next(&it->readers[adv[k]]);
    if (rc < 0) {
      free(adv);
      return rc;
    }
    if (rc == 0) heap_push(it, adv[k]);
  }
  free(adv);
  return 0;
}

static void merging_close(merging_iter* it) {
  for (int i = 0; i < it->nreaders; i++)
    if (it->readers[i].has_head) free(it->readers[i].head.values);
  free(it->readers);
  free(it->heap);
}

/* ---- outputs ---------------------------------------------------------------- */
typedef struct {
  bytebuf term_bytes, term_off, val_off, val_bytes, post, post_off, min_term, max_term;
} out_bufs;

static void out_bufs_free(out_bufs* o) {
  free(o->term_bytes.p);
  free(o->term_off.p);
  free(o->val_off.p);
  free(o->val_bytes.p);
  free(o->post.p);
  free(o->post_off.p);
  free(o->min_term.p);
  free(o->max_term.p);
  free(o);
}

static int view_check(const ii2_seg_view* s) {
  if (s->n_terms && (!s->term_off || (!s->term_bytes && s->term_off[s->n_terms]))) return -1;
  if (s->mode == II2_SEG_DECODED) return (s->n_terms && !s->post_off) ? -1 : 0;
  if (s->mode == II2_SEG_VAL || s->mode == II2_SEG_DIRECT) return (s->n_terms && !s->val_off) ? -1 : 0;
  return -1;
}

/* Shard.Merge hot loop, shard.go:158-212 + Writer.Append, file/writer.go:32-59 */
int orc_merge(const ii2_seg_view* segs, int nseg, const uint32_t* removed, uint64_t nrem,
              uint32_t flags, ii2_merge_out* out) {
  memset(out, 0, sizeof(*out));
  for (int s = 0; s < nseg; s++)
    if (view_check(&segs[s])) return II2_ERR_INVALID;
  out_bufs* ob = (out_bufs*)calloc(1, sizeof(out_bufs));
  if (!ob) return II2_ERR_NOMEM;
  merging_iter it;
  int rc = merging_open(&it, segs, nseg, NULL, 0, 0, NULL, 0, 0);
  uint64_t values_offset = 0, terms_count = 0, post_total = 0;
  uint32_t term_total = 0;
  uint32_t* enc = NULL;
  size_t enc_cap = 0;
  while (rc == 0) {
    term_values tv;
    rc = merging_next(&it, &tv);
    if (rc != 0) break;
    out->terms_merged++;
    out->postings_in += tv.nvalues; /* post-union count; refined below */
    /* minTerm/maxTerm BEFORE filtering, shard.go:176-179 */
    if (!out->has_minmax) {
      out->has_minmax = 1;
      ob->min_term.n = 0;
      bb_append(&ob->min_term, tv.term, tv.term_len);
    }
    ob->max_term.n = 0;
    bb_append(&ob->max_term, tv.term, tv.term_len);
    /* removed filter, shard.go:181-190 */
    size_t k = 0;
    for (size_t i = 0; i < tv.nvalues; i++) {
      uint32_t v = tv.values[i];
      uint64_t lo = 0, hi = nrem;
      while (lo < hi) { /* slices.BinarySearch */
        uint64_t mid = lo + (hi - lo) / 2;
        if (removed[mid] < v)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo < nrem && removed[lo] == v) continue;
      tv.values[k++] = v;
    }
    tv.nvalues = k;
    if (k == 0) { /* shard.go:192-194 */
      free(tv.values);
      continue;
    }
    /* Writer.Append: FST gets (term, valuesOffset), file/writer.go:43 */
    if (terms_count == 0) bb_append(&ob->term_off, &term_total, 4);
    bb_append(&ob->term_bytes, tv.term, tv.term_len);
    term_total += tv.term_len;
    bb_append(&ob->term_off, &term_total, 4);
    bb_append(&ob->val_off, &values_offset, 8);
    size_t bound = orc_intcomp_bound(k);
    if (bound > enc_cap) {
      enc_cap = bound * 2;
      enc = (uint32_t*)realloc(enc, enc_cap * sizeof(uint32_t));
    }
    size_t nw = orc_intcomp_encode(tv.values, k, enc); /* :49 */
    bb_append(&ob->val_bytes, enc, nw * 4);             /* :52 little-endian */
    values_offset += 4 * (uint64_t)nw;                  /* :56 */
    if (flags & II2_MERGE_WANT_DECODED) {
      if (terms_count == 0) bb_append(&ob->post_off, &post_total, 8);
      bb_append(&ob->post, tv.values, k * 4);
      uint64_t np = post_total + k;
      bb_append(&ob->post_off, &np, 8);
    }
    post_total += k;
    terms_count++;
    free(tv.values);
  }
  merging_close(&it);
  free(enc);
  if (rc < 0) {
    out_bufs_free(ob);
    memset(out, 0, sizeof(*out));
    return rc;
  }
  /* Σ input list lengths, independent of the union */
  out->postings_in = 0;
  for (int s = 0; s < nseg; s++) {
    const ii2_seg_view* v = &segs[s];
    if (v->mode == II2_SEG_DECODED)
      out->postings_in += v->n_terms ? v->post_off[v->n_terms] - v->post_off[0] : 0;
    else if (v->mode == II2_SEG_DIRECT)
      out->postings_in += v->n_terms;
    else
      for (uint64_t i = 0; i < v->n_terms; i++) {
        uint64_t end = i + 1 < v->n_terms ? v->val_off[i + 1] : v->val_size;
        size_t c = orc_intcomp_count((const uint32_t*)(v->val_bytes + v->val_off[i]),
                                     (size_t)((end - v->val_off[i]) / 4));
        out->postings_in += c == (size_t)-1 ? 0 : c;
      }
  }
  if (terms_count == 0) {
    uint32_t z = 0;
    bb_append(&ob->term_off, &z, 4);
    if (flags & II2_MERGE_WANT_DECODED) {
      uint64_t z8 = 0;
      bb_append(&ob->post_off, &z8, 8);
    }
  }
  out->terms_count = terms_count;
  out->term_bytes = ob->term_bytes.p;
  out->term_off = (uint32_t*)ob->term_off.p;
  out->val_off = (uint64_t*)ob->val_off.p;
  out->val_bytes = ob->val_bytes.p;
  out->val_size = values_offset;
  out->min_term = ob->min_term.p;
  out->min_term_len = (uint32_t)ob->min_term.n;
  out->max_term = ob->max_term.p;
  out->max_term_len = (uint32_t)ob->max_term.n;
  out->post = (uint32_t*)ob->post.p;
  out->post_off = (uint64_t*)ob->post_off.p;
  out->postings_out = post_total;
  out->_owner = ob;
  return II2_OK;
}

void orc_merge_out_free(ii2_merge_out* out) {
  if (out && out->_owner) out_bufs_free((out_bufs*)out->_owner);
  if (out) memset(out, 0, sizeof(*out));
}

/* Shard.Read, shard.go:72-75 -> makeIterator(all segments, min, max), drained like
 * go_iterators.ToSlice.  With removed != NULL the Merge-style filter is applied after
 * the union (benchmark composition, survey Q2). */
int orc_read_range(const ii2_seg_view* segs, int nseg, const uint8_t* min, size_t minlen,
                   const uint8_t* max, size_t maxlen, const uint32_t* removed, uint64_t nrem,
                   ii2_read_out* out) {
  memset(out, 0, sizeof(*out));
  for (int s = 0; s < nseg; s++)
    if (view_check(&segs[s])) return II2_ERR_INVALID;
  out_bufs* ob = (out_bufs*)calloc(1, sizeof(out_bufs));
  if (!ob) return II2_ERR_NOMEM;
  merging_iter it;
  int rc = merging_open(&it, segs, nseg, min, minlen, min != NULL, max, maxlen, max != NULL);
  uint32_t term_total = 0;
  uint64_t post_total = 0, nterms = 0;
  bb_append(&ob->term_off, &term_total, 4);
  bb_append(&ob->post_off, &post_total, 8);
  while (rc == 0) {
    term_values tv;
    rc = merging_next(&it, &tv);
    if (rc != 0) break;
    if (removed) {
      size_t k = 0;
      for (size_t i = 0; i < tv.nvalues; i++) {
        uint32_t v = tv.values[i];
        uint64_t lo = 0, hi = nrem;
        while (lo < hi) {
          uint64_t mid = lo + (hi - lo) / 2;
          if (removed[mid] < v)
            lo = mid + 1;
          else
            hi = mid;
        }
        if (lo < nrem && removed[lo] == v) continue;
        tv.values[k++] = v;
      }
      tv.nvalues = k;
      if (k == 0) {
        free(tv.values);
        continue;
      }
    }
    bb_append(&ob->term_bytes, tv.term, tv.term_len);
    term_total += tv.term_len;
    bb_append(&ob->term_off, &term_total, 4);
    bb_append(&ob->post, tv.values, tv.nvalues * 4);
    post_total += tv.nvalues;
    bb_append(&ob->post_off, &post_total, 8);
    nterms++;
    free(tv.values);
  }
  merging_close(&it);
  if (rc < 0) {
    out_bufs_free(ob);
    return rc;
  }
  out->n_terms = nterms;
  out->term_bytes = ob->term_bytes.p;
  out->term_off = (uint32_t*)ob->term_off.p;
  out->post = (uint32_t*)ob->post.p;
  out->post_off = (uint64_t*)ob->post_off.p;
  out->_owner = ob;
  return II2_OK;
}

void orc_read_out_free(ii2_read_out* out) {
  if (out && out->_owner) out_bufs_free((out_bufs*)out->_owner);
  if (out) memset(out, 0, sizeof(*out));
}

/* shardKey, shard.go:362-378 (numeric value; the reference formats it "%04d") */
uint32_t orc_shard_key(const uint8_t* term, size_t len) {
  uint16_t key = 0;
  if (len >= 2) key = (uint16_t)(((uint16_t)term[0] << 8) + term[1]);
  return (uint32_t)(key >> 6);
}

void orc_free(void* p) { free(p); }
