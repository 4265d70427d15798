/*
 * ii2_oracle.h — CPU oracle for the hot path of lezhnev74/inverted_index_2.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product (libii2.so) never links or calls it.
 *
 * It is a plain-C restatement of the reference's ALGORITHM for the path
 * (pull-style readers, k-way merge with pairwise append+sort+compact unions,
 * per-value binary search in the removed list, one codec call per term), each
 * function citing the reference file:line it follows.  It shares the flat
 * view/out structs of include/ii2.h so the same inputs can be handed to both.
 *
 * PARITY STATUS
 *   logical results (term -> values, counts, min/max, drop rules):
 *       PINNED against every known-answer vector in the reference's own tests
 *       (tests/golden/reference_vectors.json, tests/test_oracle_golden.py).
 *   bytes of <key>_val (ronanh/intcomp v1.1.0, go.mod:10) and of
 *   Bitmask.Put (RoaringBitmap/roaring v1.9.4, go.mod:6):
 *       PARITY UNPINNED.  Those modules are not vendored under /root/reference,
 *       no Go toolchain exists in this image, and the reference's tests hold
 *       no byte-level vectors.  intcomp_ref.c / roaring_ref.c restate the
 *       published formats (see their headers for per-item confidence).
 */
#ifndef II2_ORACLE_H
#define II2_ORACLE_H

#include "../include/ii2.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- intcomp restatement (intcomp_ref.c) -------------------------------- */
/* Worst-case output words for n input values. */
size_t orc_intcomp_bound(size_t n);
/* intcomp.CompressUint32(in, nil) — call site file/writer.go:49. Returns words written. */
size_t orc_intcomp_encode(const uint32_t* in, size_t n, uint32_t* out);
/* Number of values a stream decodes to (reads headers only); (size_t)-1 if corrupt. */
size_t orc_intcomp_count(const uint32_t* words, size_t nwords);
/* intcomp.UncompressUint32(in, nil) — call site file/reader.go:100.
 * Returns values written, (size_t)-1 if corrupt or cap too small. */
size_t orc_intcomp_decode(const uint32_t* words, size_t nwords, uint32_t* out, size_t cap);
/* batched: list i = in[off[i]..off[i+1]) -> out words [word_off[i]..word_off[i+1]) */
uint64_t orc_intcomp_encode_batch(const uint32_t* in, const uint64_t* off, uint64_t nlists,
                                  uint32_t* out, uint64_t* word_off);
uint64_t orc_intcomp_count_batch(const uint32_t* words, const uint64_t* word_off, uint64_t nlists,
                                 uint64_t* out_off);
int orc_intcomp_decode_batch(const uint32_t* words, const uint64_t* word_off, uint64_t nlists,
                             const uint64_t* out_off, uint32_t* out);

/* ---- roaring + Bitmask restatement (roaring_ref.c) ----------------------- */
typedef struct orc_bitmask orc_bitmask;
orc_bitmask* orc_bitmask_new(const uint32_t* init, uint64_t n);
void orc_bitmask_free(orc_bitmask* bm);
uint64_t orc_bitmask_len(const orc_bitmask* bm);
const uint32_t* orc_bitmask_values(const orc_bitmask* bm);
/* fast != 0 answers slices.Index through a hash map (same result, not the
 * reference's O(L*D) shape) so large cases can be checked. */
int orc_bitmask_put(orc_bitmask* bm, const uint32_t* vals, uint64_t n, int fast, uint8_t** bytes,
                    uint64_t* nbytes);
int orc_bitmask_get(const orc_bitmask* bm, const uint8_t* enc, uint64_t nenc, uint32_t** vals,
                    uint64_t* n);

/* ---- merge / read restatement (merge_ref.c) ------------------------------ */
int orc_merge(const ii2_seg_view* segs, int nseg, const uint32_t* removed_sorted, uint64_t nrem,
              uint32_t flags, ii2_merge_out* out);
void orc_merge_out_free(ii2_merge_out* out);
int orc_read_range(const ii2_seg_view* segs, int nseg, const uint8_t* min, size_t minlen,
                   const uint8_t* max, size_t maxlen, const uint32_t* removed_sorted,
                   uint64_t nrem, ii2_read_out* out);
void orc_read_out_free(ii2_read_out* out);
uint32_t orc_shard_key(const uint8_t* term, size_t len);
void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
