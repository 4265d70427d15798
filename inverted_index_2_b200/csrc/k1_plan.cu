// k1_plan.cu — K1: splitters for the k-way merge of the sorted term dictionaries.
//
// Replaces the head selection of go-iterators' MergingIterator over per-segment readers
// (shard.go:253-278, ordering = file.CompareTermValues = bytes.Compare, file/types.go:24-26)
// by a merge-path style partition of the OUTPUT term space:
//
//   k1_sample_keys     every segment contributes evenly spaced sample terms in proportion to
//                      its size (about one per 640 instances overall); 16-byte key windows.
//   k1_rank_partition  one warp per sample x, lanes over segments: (a) binary search among the
//                      segment's own samples -> how many are smaller -> the rank of x in the
//                      merged sample order is the sum over segments (a k-way merge by ranking,
//                      no sort); (b) the two neighbouring samples bound the lower_bound of x in
//                      the full segment to ~n/m terms, finished by a short binary search.
//                      Row x of `part` = lower_bound of x in every segment.
//   k1_bucket_stats    per bucket: instances, input postings (from the posting offsets),
//                      common prefix length; then one scan -> bucket bases.
//
// Integer/byte work; every probe is an L2 hit after the first touch (samples and offsets of
// 64 segments are a few MB).
#include "keys.cuh"
#include "plan.cuh"

namespace ii2 {

constexpr uint32_t kInstancesPerBucket = 640;  // target; K12 tiles hold 1024
constexpr uint32_t kMaxBuckets = 1u << 20;

struct SampleArrays {
  uint64_t* hi;
  uint64_t* lo;
  uint64_t* ptr;   // first byte of the sample term
  uint32_t* len;
  uint32_t* idx;   // term index inside its segment
  uint32_t* seg;
};

__global__ void __launch_bounds__(256)
k1_sample_keys(const SegDesc* __restrict__ segs, int k, const uint32_t* __restrict__ sbase,
               uint32_t S, SampleArrays sa) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= S) return;
  int lo = 0, hi = k;  // segment of sample x: last s with sbase[s] <= x
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sbase[mid + 1] <= x)
      lo = mid + 1;
    else
      hi = mid;
  }
  const int s = lo;
  const SegDesc sd = segs[s];
  const uint32_t m = sbase[s + 1] - sbase[s], j = x - sbase[s], n = sd.hi - sd.lo;
  const uint32_t idx = sd.lo + (uint32_t)(((uint64_t)(j + 1) * n) / (m + 1));
  const KeyedTerm t = keyed_term(sd, idx);
  sa.hi[x] = t.hi;
  sa.lo[x] = t.lo;
  sa.ptr[x] = reinterpret_cast<uint64_t>(t.p);
  sa.len[x] = t.len;
  sa.idx[x] = idx;
  sa.seg[x] = (uint32_t)s;
}

__device__ __forceinline__ KeyedTerm sample_term(const SampleArrays& sa, uint32_t x) {
  KeyedTerm t;
  t.hi = sa.hi[x];
  t.lo = sa.lo[x];
  t.p = reinterpret_cast<const uint8_t*>(sa.ptr[x]);
  t.len = sa.len[x];
  return t;
}

__global__ void __launch_bounds__(256)
k1_rank_partition(const SegDesc* __restrict__ segs, int k, const uint32_t* __restrict__ sbase,
                  uint32_t S, SampleArrays sa, uint32_t* __restrict__ part,
                  uint32_t* __restrict__ row_of) {
  const uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = lane_id();
  if (x >= S + 2) return;
  if (x >= S) {  // row S = window starts, row S+1 = window ends
    for (int s = lane; s < k; s += 32) part[(uint64_t)x * k + s] = x == S ? segs[s].lo : segs[s].hi;
    if (lane == 0) row_of[x == S ? 0 : S + 1] = x;
    return;
  }
  const KeyedTerm tx = sample_term(sa, x);
  const uint32_t sx = sa.seg[x];
  uint32_t racc = 0;
  for (int s = lane; s < k; s += 32) {
    const uint32_t b0 = sbase[s], m = sbase[s + 1] - b0;
    uint32_t lb;
    if ((uint32_t)s == sx) {
      racc += x - b0;
      lb = sa.idx[x];
    } else {
      uint32_t lo = 0, hi = m;  // samples of s strictly below x
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keyed_compare(sample_term(sa, b0 + mid), tx) < 0)
          lo = mid + 1;
        else
          hi = mid;
      }
      const uint32_t q = lo;
      const SegDesc sd = segs[s];
      const uint32_t wlo = q > 0 ? sa.idx[b0 + q - 1] + 1 : sd.lo;
      uint32_t whi = sd.hi;
      bool equal = false;
      if (q < m) {
        whi = sa.idx[b0 + q];
        equal = keyed_compare(sample_term(sa, b0 + q), tx) == 0;
      }
      // ties between equal sample terms are ordered by segment
      racc += q + ((equal && (uint32_t)s < sx) ? 1u : 0u);
      lb = equal ? whi : keyed_lower_bound(sd, wlo, whi, tx);
    }
    part[(uint64_t)x * k + s] = lb;
  }
  racc = warp_sum(racc);
  if (lane == 0) row_of[racc + 1] = x;
}

// One warp per bucket.  raw[0][b] = instances, raw[1][b] = input postings.
__global__ void __launch_bounds__(256)
k1_bucket_stats(const SegDesc* __restrict__ segs, int k, uint32_t S, SampleArrays sa,
                const uint32_t* __restrict__ part, const uint32_t* __restrict__ row_of,
                uint64_t* __restrict__ raw, uint32_t* __restrict__ bk_cpl) {
  const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t B = S + 1;
  if (b > B) return;
  const unsigned lane = lane_id();
  if (b == B) {
    if (lane == 0) raw[B] = raw[(uint64_t)(B + 1) + B] = 0;
    return;
  }
  const uint32_t r0 = row_of[b], r1 = row_of[b + 1];
  uint64_t w = 0, p = 0;
  for (int s = lane; s < k; s += 32) {
    const uint32_t a = part[(uint64_t)r0 * k + s], e = part[(uint64_t)r1 * k + s];
    w += e - a;
    if (e > a) p += __ldg(segs[s].poff + e) - __ldg(segs[s].poff + a);
  }
  w = warp_sum(w);
  p = warp_sum(p);
  if (lane == 0) {
    raw[b] = w;
    raw[(uint64_t)(B + 1) + b] = p;
    uint32_t c = 0;
    if (r0 < S && r1 < S) {  // both delimiting splitters exist
      const uint8_t* x = reinterpret_cast<const uint8_t*>(sa.ptr[r0]);
      const uint8_t* y = reinterpret_cast<const uint8_t*>(sa.ptr[r1]);
      const uint32_t m = sa.len[r0] < sa.len[r1] ? sa.len[r0] : sa.len[r1];
      while (c < m && x[c] == y[c]) c++;
    }
    bk_cpl[b] = c;
  }
}

int k1_build_plan(MergePlan& plan, const SegDesc* h_segs, uint32_t* sbase, cudaStream_t s) {
  const int k = plan.k;
  const uint32_t N = plan.n_total;
  if (k > kMaxSegs) {
    set_last_error("k1: %d segments in one pass (max %d)", k, kMaxSegs);
    return II2_ERR_UNSUPPORTED;
  }
  // samples per segment, proportional to its window
  uint64_t want = N / kInstancesPerBucket;
  if (want > kMaxBuckets - 1) want = kMaxBuckets - 1;
  sbase[0] = 0;
  for (int i = 0; i < k; i++) {
    const uint64_t n = h_segs[i].hi - h_segs[i].lo;
    uint64_t m = N ? want * n / N : 0;
    if (m + 1 > n) m = n ? n - 1 : 0;
    sbase[i + 1] = sbase[i] + (uint32_t)m;
  }
  const uint32_t S = sbase[k], B = S + 1;
  plan.n_samples = S;
  plan.n_buckets = B;
  ProfScope scope("k1_plan", s);
  II2_TRY(plan.part.alloc_scratch((size_t)(S + 2) * k, s));
  II2_TRY(plan.row_of.alloc_scratch(B + 1, s));
  II2_TRY(plan.bk_cpl.alloc_scratch(B, s));
  II2_TRY(plan.bk_WP.alloc_scratch(2 * (size_t)(B + 1), s));
  II2_TRY(plan.totals.alloc_scratch(2, s));
  DevBuf<uint32_t> d_sbase, d_u32;
  DevBuf<uint64_t> d_u64;
  II2_TRY(d_sbase.alloc_scratch(k + 1, s));
  II2_TRY(d_u64.alloc_scratch(3 * (size_t)(S ? S : 1), s));
  II2_TRY(d_u32.alloc_scratch(3 * (size_t)(S ? S : 1), s));
  II2_CUDA_TRY(cudaMemcpyAsync(d_sbase.p, sbase, (k + 1) * 4, cudaMemcpyHostToDevice, s));
  const size_t Sx = S ? S : 1;
  SampleArrays sa{d_u64.p, d_u64.p + Sx, d_u64.p + 2 * Sx, d_u32.p, d_u32.p + Sx, d_u32.p + 2 * Sx};
  if (S) {
    k1_sample_keys<<<div_up(S, 256), 256, 0, s>>>(plan.segs, k, d_sbase.p, S, sa);
    II2_LAUNCHED();
  }
  k1_rank_partition<<<div_up((uint64_t)(S + 2) * 32, 256), 256, 0, s>>>(plan.segs, k, d_sbase.p, S,
                                                                       sa, plan.part.p,
                                                                       plan.row_of.p);
  II2_LAUNCHED();
  k1_bucket_stats<<<div_up((uint64_t)(B + 1) * 32, 256), 256, 0, s>>>(
      plan.segs, k, S, sa, plan.part.p, plan.row_of.p, plan.bk_WP.p, plan.bk_cpl.p);
  II2_LAUNCHED();
  II2_TRY(exclusive_scan_multi_u64(plan.bk_WP.p, plan.bk_WP.p, B + 1, 2, plan.totals.p, s));
  return II2_OK;
}

}  // namespace ii2
