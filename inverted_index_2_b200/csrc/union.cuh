// union.cuh — outputs of K2 (posting union) and inputs of the emit kernel.
#pragma once
#include "plan.cuh"

namespace ii2 {

struct UnionOut {
  DevBuf<uint32_t> tmp_post;  // [N_in] union results; group at head p starts at g_off[p]
  DevBuf<uint32_t> g_cnt;     // [N_T] at heads: values left after union + removed filter
  DevBuf<uint32_t> g_enc;     // [N_T] at heads: intcomp words of that list (if encoding)
  DevBuf<uint64_t> g_off;     // [N_T] at heads: offset into tmp_post
  // [4][B+1] per bucket {surviving terms, their term bytes, postings out, encoded words}
  DevBuf<uint64_t> bk_raw;
  DevBuf<uint64_t> bk_out;    // exclusive prefixes of bk_raw
  DevBuf<uint64_t> totals;    // [4] device
  uint64_t h_totals[4] = {0, 0, 0, 0};
  bool keep_empty = false;    // reads without a filter keep terms whose list is empty
};

// K2: per-term union + dedup of uint32 posting lists with the removed filter in the same pass
// (file.MergeTermValues file/types.go:14-22 for terms with >= 2 sources, pass-through for
// single-source terms; filter shard.go:181-190).  Synchronises the stream once to learn
// the output sizes.
// n_in_hint: Σ input postings if the host already knows it (full-window merges), else 0.
// keep_empty: terms left with no values are kept (plain reads) instead of dropped (merge,
// shard.go:192-194).
int k2_union(const MergePlan& plan, const RemovedSet& rem, bool want_enc, bool keep_empty,
             uint64_t n_in_hint, UnionOut& u, cudaStream_t s);

struct EmitOut {
  DevBuf<uint8_t> term_bytes;
  DevBuf<uint32_t> term_off;  // [T+1]
  DevBuf<uint32_t> post;      // [P]    (if want_decoded)
  DevBuf<uint64_t> post_off;  // [T+1]  (if want_decoded)
  DevBuf<uint32_t> val_words; // [E]    (if want_enc)
  DevBuf<uint64_t> val_off;   // [T]    byte offsets (if want_enc)
};

// Emit: surviving terms (non-empty after the filter, shard.go:192-194) with compact term
// bytes/offsets, decoded postings and/or the intcomp-encoded `_val` stream with running byte
// offsets (Writer.Append, file/writer.go:43-56).
int k6_emit(const MergePlan& plan, const UnionOut& u, bool want_decoded, bool want_enc,
            EmitOut& out, cudaStream_t s);

}  // namespace ii2
