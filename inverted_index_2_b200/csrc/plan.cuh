// plan.cuh — data handed from K1 (term-dictionary merge) to K2 (posting union) and the
// emit kernel.  All buffers are device memory owned by the plan.
//
// Vocabulary
//   instance  one (segment, term) pair inside the call's windows; global instance id
//             g = segs[s].base + (idx - segs[s].lo); N_T instances in total.
//   position  index p in [0, N_T) of an instance in MERGED order: ascending term, all
//             instances of one term contiguous (a "group"); the first is the head.
//   bucket    contiguous range of positions produced by one K1 CTA; buckets are delimited
//             by sorted sample terms ("splitters"), so no term straddles two buckets.
#pragma once
#include "runtime.cuh"

namespace ii2 {

struct MergePlan {
  int k = 0;                      // segments
  uint32_t n_total = 0;           // N_T
  uint32_t n_buckets = 0;         // B (= samples + 1)
  const SegDesc* segs = nullptr;  // [k] device, windows filled in
  DevBuf<uint32_t> part;          // [(B+1) * k] per-bucket per-segment lower bounds
  DevBuf<uint32_t> bk_pos;        // [B+1] first position of each bucket
  DevBuf<uint32_t> bk_cpl;        // [B]   common prefix length of all terms in the bucket
  DevBuf<uint32_t> ord_inst;      // [N_T] instance id at each position
  DevBuf<uint64_t> src_ptr;       // [N_T] device address of the instance's posting list
  DevBuf<uint32_t> src_len;       // [N_T] its length
  DevBuf<uint16_t> gsz;           // [N_T] group size at head positions, 0 elsewhere
  DevBuf<uint64_t> bk_PD;         // [2][B+1] postings / distinct terms per bucket -> exclusive
                                  //          prefixes after the scan
  uint64_t* bk_P() const { return bk_PD.p; }
  uint64_t* bk_D() const { return bk_PD.p + (n_buckets + 1); }
  DevBuf<uint64_t> totals;        // [2]   {Σ postings in, Σ distinct terms}
};

// K1: k-way merge of the segments' term dictionaries (go-iterators MergingIterator built at
// shard.go:267 with file.CompareTermValues).  plan.k / n_total / segs must be set.
int k1_build_plan(MergePlan& plan, cudaStream_t s);

}  // namespace ii2
