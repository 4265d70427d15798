// k2_posting_union.cu — K2: per-term union + dedup of uint32 posting lists, removed filter in
// the same pass.
//
// Replaces file.MergeTermValues (file/types.go:14-22: append + slices.Sort + slices.Compact,
// applied pairwise by the merging iterator) and the removed filter of the merge loop
// (shard.go:181-190).  Semantics kept bit-exact:
//   - a term present in ONE segment passes through untouched — not sorted, not deduped
//     (survey Q4); a term present in >= 2 segments becomes the sorted-unique union;
//   - the filter runs AFTER the union; removed = membership in the sorted removed list.
//
// Work decomposition: one CTA per K1 bucket walks its positions in chunks of 256; the heads
// found in a chunk are handed to warps.  Per group, by total input length L:
//   1 source          warp streams the list through the filter (ballot compaction)
//   L <= 1024         warp: gather -> bitonic sort in its shared-memory slice -> dedup + filter
//                     by ballot/popc compaction
//   L <= 8192         whole CTA: same in the full 32 KB shared buffer
//   larger            deferred to the multi-CTA global-memory path below
// Results go to tmp_post at the group's input-prefix offset (an upper bound of its output
// size); the emit kernel compacts.  Encoded size of every list is computed here while the list
// is still in shared memory, so `_val` offsets are ready after one bucket-level scan.
#include <algorithm>

#include "intcomp.cuh"
#include "union.cuh"

namespace ii2 {

constexpr int K2_THREADS = 256;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr uint32_t K2_WCAP = 1024;   // values a warp sorts in its scratch slice
constexpr uint32_t K2_CCAP = 8192;   // values the CTA sorts in shared memory
constexpr int K2_DEFER = 64;         // CTA-level groups remembered per bucket
constexpr uint32_t K2_PENDING = 0xFFFFFFFFu;

struct K2Args {
  const SegDesc* segs;
  int k;
  const uint32_t* bk_pos;
  const uint64_t* bk_Pbase;
  const uint32_t* ord_inst;
  const uint64_t* src_ptr;
  const uint32_t* src_len;
  const uint16_t* gsz;
  RemovedSet rem;
  int want_enc;
  int keep_empty;
  uint32_t* tmp_post;
  uint32_t* g_cnt;
  uint32_t* g_enc;
  uint64_t* g_off;
  uint64_t* bk_out;  // [4][nb1]
  uint32_t nb1;      // B + 1
  uint32_t* large_pos;   // deferred heads
  uint64_t* large_len;   // their total input length
  uint32_t* n_large;
};

__device__ __forceinline__ uint32_t term_len_of(const SegDesc* segs, int k, uint32_t inst) {
  int s;
  uint32_t idx;
  locate_instance(segs, k, inst, s, idx);
  return segs[s].toff[idx + 1] - segs[s].toff[idx];
}

// record a finished group (one thread)
__device__ __forceinline__ void finish_group(const K2Args& a, unsigned long long* acc, uint32_t hp,
                                             uint32_t outn, uint32_t enc, uint64_t off) {
  a.g_cnt[hp] = outn;
  a.g_enc[hp] = enc;
  a.g_off[hp] = off;
  if (outn || a.keep_empty) {
    atomicAdd(&acc[0], 1ull);
    atomicAdd(&acc[1], (unsigned long long)term_len_of(a.segs, a.k, a.ord_inst[hp]));
    atomicAdd(&acc[2], (unsigned long long)outn);
    atomicAdd(&acc[3], (unsigned long long)enc);
  }
}

// ---- one warp: group at head position hp with c sources, tmp offset off --------------------
__device__ __forceinline__ void union_group_warp(const K2Args& a, uint32_t hp, uint32_t c,
                                                 uint64_t off, uint32_t* ws,
                                                 unsigned long long* acc, uint32_t* s_ndefer,
                                                 uint32_t* s_defer, uint64_t* s_defer_off) {
  const unsigned lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  uint64_t L = 0;
  for (uint32_t i = lane; i < c; i += 32) L += a.src_len[hp + i];
  L = warp_sum(L);
  uint32_t* dst = a.tmp_post + off;

  if (c == 1) {  // pass-through: keep order and duplicates, only filter
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.src_ptr[hp]);
    uint64_t outn = 0;
    for (uint64_t e0 = 0; e0 < L; e0 += 32) {
      uint64_t e = e0 + lane;
      bool valid = e < L;
      uint32_t v = valid ? __ldg(src + e) : 0u;
      bool keep = valid && !is_removed(a.rem, v);
      unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) dst[outn + __popc(bal & lt)] = v;
      outn += __popc(bal);
    }
    __syncwarp();
    uint32_t enc = a.want_enc ? intcomp::enc_size_warp(dst, (uint32_t)outn) : 0u;
    if (lane == 0) finish_group(a, acc, hp, (uint32_t)outn, enc, off);
    return;
  }

  if (L > K2_WCAP) {  // too big for a warp slice: defer
    if (lane == 0) {
      bool to_large = L > K2_CCAP;
      if (!to_large) {
        uint32_t slot = atomicAdd(s_ndefer, 1u);
        if (slot < (uint32_t)K2_DEFER) {
          s_defer[slot] = hp;
          s_defer_off[slot] = off;
        } else {
          to_large = true;
        }
      }
      if (to_large) {
        uint32_t idx = atomicAdd(a.n_large, 1u);
        a.large_pos[idx] = hp;
        a.large_len[idx] = L;
        a.g_off[hp] = off;
        a.g_cnt[hp] = K2_PENDING;
      }
    }
    return;
  }

  // gather the c lists into the warp's slice (load-balanced over the flattened values)
  uint32_t filled = 0;
  for (uint32_t i0 = 0; i0 < c; i0 += 32) {
    uint32_t i = i0 + lane;
    uint32_t li = i < c ? a.src_len[hp + i] : 0u;
    uint64_t pi = i < c ? a.src_ptr[hp + i] : 0ull;
    uint32_t inc = warp_inclusive_scan(li);
    uint32_t ex = inc - li;
    uint32_t T = __shfl_sync(0xffffffffu, inc, 31);
    for (uint32_t e0 = 0; e0 < T; e0 += 32) {
      uint32_t e = e0 + lane;
      uint32_t m = 0;  // last lane whose exclusive offset is <= e
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        uint32_t cand = m + step;
        uint32_t exc = __shfl_sync(0xffffffffu, ex, cand & 31u);
        if (cand < 32u && exc <= e) m = cand;
      }
      uint64_t pm = __shfl_sync(0xffffffffu, pi, m);
      uint32_t exm = __shfl_sync(0xffffffffu, ex, m);
      if (e < T) ws[filled + e] = __ldg(reinterpret_cast<const uint32_t*>(pm) + (e - exm));
    }
    filled += T;
  }
  __syncwarp();
  const uint32_t n = (uint32_t)L;
  bitonic_sort_any(ws, n, lane, 32u, [](uint32_t x, uint32_t y) { return x < y; },
                   [] { __syncwarp(); });
  __syncwarp();
  // dedup (slices.Compact) + removed filter, compacted in place
  uint32_t outn = 0;
  for (uint32_t e0 = 0; e0 < n; e0 += 32) {
    uint32_t e = e0 + lane;
    bool valid = e < n;
    uint32_t v = valid ? ws[e] : 0u;
    bool keep = valid && (e == 0 || ws[e - 1] != v) && !is_removed(a.rem, v);
    __syncwarp();
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) ws[outn + __popc(bal & lt)] = v;
    outn += __popc(bal);
    __syncwarp();
  }
  for (uint32_t e = lane; e < outn; e += 32) dst[e] = ws[e];
  uint32_t enc = a.want_enc ? intcomp::enc_size_warp(ws, outn) : 0u;
  if (lane == 0) finish_group(a, acc, hp, outn, enc, off);
  __syncwarp();
}

// ---- whole CTA: group with K2_WCAP < L <= K2_CCAP ---------------------------------------------
__device__ __forceinline__ void union_group_cta(const K2Args& a, uint32_t hp, uint64_t off,
                                                uint32_t* scratch, uint32_t* s_moff,
                                                uint32_t* s_ws32, uint32_t* s_bc,
                                                unsigned long long* acc) {
  const uint32_t tid = threadIdx.x;
  const uint32_t c = a.gsz[hp];
  uint32_t run = 0;
  for (uint32_t base = 0; base < c; base += K2_THREADS) {
    uint32_t i = base + tid;
    uint32_t li = i < c ? a.src_len[hp + i] : 0u;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(li, s_ws32, tot);
    if (i < c) s_moff[i] = run + ex;
    run += tot;
  }
  const uint32_t n = run;
  if (tid == 0) s_moff[c] = n;
  __syncthreads();
  for (uint32_t e = tid; e < n; e += K2_THREADS) {
    uint32_t lo = 0, hi = c;  // first m with moff[m+1] > e
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (s_moff[mid + 1] <= e)
        lo = mid + 1;
      else
        hi = mid;
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.src_ptr[hp + lo]);
    scratch[e] = __ldg(src + (e - s_moff[lo]));
  }
  __syncthreads();
  bitonic_sort_any(scratch, n, tid, (uint32_t)K2_THREADS,
                   [](uint32_t x, uint32_t y) { return x < y; }, [] { __syncthreads(); });
  __syncthreads();
  uint32_t outn = 0;
  for (uint32_t e0 = 0; e0 < n; e0 += K2_THREADS) {
    uint32_t e = e0 + tid;
    bool valid = e < n;
    uint32_t v = valid ? scratch[e] : 0u;
    uint32_t keep = (valid && (e == 0 || scratch[e - 1] != v) && !is_removed(a.rem, v)) ? 1u : 0u;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(keep, s_ws32, tot);  // syncs: all reads precede writes
    if (keep) scratch[outn + ex] = v;
    outn += tot;
    __syncthreads();
  }
  uint32_t* dst = a.tmp_post + off;
  for (uint32_t e = tid; e < outn; e += K2_THREADS) dst[e] = scratch[e];
  if (warp_id() == 0) {
    uint32_t enc = a.want_enc ? intcomp::enc_size_warp(scratch, outn) : 0u;
    if (lane_id() == 0) finish_group(a, acc, hp, outn, enc, off);
  }
  (void)s_bc;
  __syncthreads();
}

__global__ void __launch_bounds__(K2_THREADS) k2_union_kernel(const K2Args a) {
  __shared__ uint32_t scratch[K2_CCAP];
  __shared__ uint64_t s_off[K2_THREADS];
  __shared__ uint16_t s_heads[K2_THREADS];
  __shared__ uint32_t s_defer[K2_DEFER];
  __shared__ uint64_t s_defer_off[K2_DEFER];
  __shared__ uint32_t s_ndefer;
  __shared__ uint64_t s_ws64[K2_WARPS + 2];
  __shared__ uint32_t s_ws32[K2_WARPS + 2];
  __shared__ unsigned long long s_acc[4];
  __shared__ uint32_t s_moff[kMaxSegs + 1];
  __shared__ uint32_t s_bc[2];

  const uint32_t tid = threadIdx.x;
  const uint32_t b = blockIdx.x;
  const uint32_t p0 = a.bk_pos[b], p1 = a.bk_pos[b + 1];
  if (tid < 4) s_acc[tid] = 0;
  if (tid == 0) s_ndefer = 0;
  __syncthreads();
  uint64_t running = a.bk_Pbase[b];
  for (uint32_t q = p0; q < p1; q += K2_THREADS) {
    const uint32_t p = q + tid;
    const bool valid = p < p1;
    const uint64_t len = valid ? a.src_len[p] : 0u;
    const uint32_t isHead = (valid && a.gsz[p] != 0) ? 1u : 0u;
    uint64_t tot;
    uint64_t ex = block_exclusive_scan(len, s_ws64, tot);
    s_off[tid] = running + ex;
    uint32_t nh;
    uint32_t hex = block_exclusive_scan(isHead, s_ws32, nh);
    if (isHead) s_heads[hex] = (uint16_t)tid;
    __syncthreads();
    for (uint32_t h = warp_id(); h < nh; h += K2_WARPS) {
      const uint32_t t = s_heads[h];
      const uint32_t hp = q + t;
      union_group_warp(a, hp, a.gsz[hp], s_off[t], scratch + warp_id() * K2_WCAP, s_acc, &s_ndefer,
                       s_defer, s_defer_off);
    }
    running += tot;
    __syncthreads();
  }
  const uint32_t nd = s_ndefer < (uint32_t)K2_DEFER ? s_ndefer : (uint32_t)K2_DEFER;
  for (uint32_t d = 0; d < nd; d++)
    union_group_cta(a, s_defer[d], s_defer_off[d], scratch, s_moff, s_ws32, s_bc, s_acc);
  __syncthreads();
  if (tid < 4) a.bk_out[(uint64_t)tid * a.nb1 + b] = s_acc[tid];
}

// ---------------------------------------------------------------- large groups (global memory)
struct LargeArgs {
  K2Args a;
  uint32_t n_large;
};

// gather: grid (x = CTAs per group, y = group)
__global__ void __launch_bounds__(256) k2_large_gather(const K2Args a) {
  __shared__ uint64_t s_moff[kMaxSegs + 1];
  __shared__ uint64_t s_ws[256 / 32 + 2];
  const uint32_t hp = a.large_pos[blockIdx.y];
  const uint32_t c = a.gsz[hp];
  uint64_t run = 0;
  for (uint32_t base = 0; base < c; base += 256) {
    uint32_t i = base + threadIdx.x;
    uint64_t li = i < c ? a.src_len[hp + i] : 0u;
    uint64_t tot;
    uint64_t ex = block_exclusive_scan(li, s_ws, tot);
    if (i < c) s_moff[i] = run + ex;
    run += tot;
  }
  if (threadIdx.x == 0) s_moff[c] = run;
  __syncthreads();
  const uint64_t n = run;
  uint32_t* dst = a.tmp_post + a.g_off[hp];
  for (uint64_t e = (uint64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (uint64_t)gridDim.x * 256) {
    uint32_t lo = 0, hi = c;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (s_moff[mid + 1] <= e)
        lo = mid + 1;
      else
        hi = mid;
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.src_ptr[hp + lo]);
    dst[e] = __ldg(src + (e - s_moff[lo]));
  }
}

constexpr uint32_t LG_TILE = 4096;

// sort every aligned LG_TILE tile of every large group in shared memory
__global__ void __launch_bounds__(512) k2_large_tile_sort(const K2Args a) {
  __shared__ uint32_t tile[LG_TILE];
  const uint32_t hp = a.large_pos[blockIdx.y];
  const uint64_t n = a.large_len[blockIdx.y];
  uint32_t* base = a.tmp_post + a.g_off[hp];
  for (uint64_t t0 = (uint64_t)blockIdx.x * LG_TILE; t0 < n; t0 += (uint64_t)gridDim.x * LG_TILE) {
    uint32_t m = (uint32_t)((n - t0) < LG_TILE ? (n - t0) : LG_TILE);
    for (uint32_t i = threadIdx.x; i < m; i += 512) tile[i] = base[t0 + i];
    __syncthreads();
    bitonic_sort_any(tile, m, threadIdx.x, 512u, [](uint32_t x, uint32_t y) { return x < y; },
                     [] { __syncthreads(); });
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < m; i += 512) base[t0 + i] = tile[i];
    __syncthreads();
  }
}

// one global stage of the direction-free bitonic network: flip (kk, j == 0) or half-cleaner j
__global__ void __launch_bounds__(256) k2_large_stage(const K2Args a, uint64_t kk, uint64_t j) {
  const uint32_t hp = a.large_pos[blockIdx.y];
  const uint64_t n = a.large_len[blockIdx.y];
  if ((kk >> 1) >= n) return;  // this group is already sorted at this block size
  uint32_t* v = a.tmp_post + a.g_off[hp];
  const uint64_t half = j ? j : (kk >> 1);
  const uint64_t limit = (n + 1) / 2 + half;
  for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < limit;
       t += (uint64_t)gridDim.x * 256) {
    uint64_t i, l;
    if (j == 0) {
      i = (t / half) * kk + (t % half);
      l = i ^ (kk - 1);
    } else {
      i = (t / j) * (j << 1) + (t % j);
      l = i + j;
    }
    if (l < n && i < n) {
      uint32_t x = v[i], y = v[l];
      if (y < x) {
        v[i] = y;
        v[l] = x;
      }
    }
  }
}

// finish block size kk inside shared memory: half-cleaners j = LG_TILE/2 .. 1
__global__ void __launch_bounds__(512) k2_large_tile_merge(const K2Args a, uint64_t kk) {
  __shared__ uint32_t tile[LG_TILE];
  const uint32_t hp = a.large_pos[blockIdx.y];
  const uint64_t n = a.large_len[blockIdx.y];
  if ((kk >> 1) >= n) return;
  uint32_t* base = a.tmp_post + a.g_off[hp];
  for (uint64_t t0 = (uint64_t)blockIdx.x * LG_TILE; t0 < n; t0 += (uint64_t)gridDim.x * LG_TILE) {
    uint32_t m = (uint32_t)((n - t0) < LG_TILE ? (n - t0) : LG_TILE);
    for (uint32_t i = threadIdx.x; i < m; i += 512) tile[i] = base[t0 + i];
    __syncthreads();
    for (uint32_t j = LG_TILE / 2; j >= 1; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < LG_TILE / 2; t += 512) {
        uint32_t i = (t / j) * (j << 1) + (t % j);
        uint32_t l = i + j;
        if (l < m) {
          uint32_t x = tile[i], y = tile[l];
          if (y < x) {
            tile[i] = y;
            tile[l] = x;
          }
        }
      }
      __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < m; i += 512) base[t0 + i] = tile[i];
    __syncthreads();
  }
}

// one CTA per large group: in-place dedup + filter, encoded size, bucket totals
__global__ void __launch_bounds__(1024) k2_large_finish(const K2Args a) {
  __shared__ uint64_t s_ws[1024 / 32 + 2];
  __shared__ uint32_t s_enc;
  const uint32_t hp = a.large_pos[blockIdx.x];
  const uint64_t n = a.large_len[blockIdx.x];
  uint32_t* v = a.tmp_post + a.g_off[hp];
  uint64_t outn = 0;
  for (uint64_t e0 = 0; e0 < n; e0 += 1024) {
    uint64_t e = e0 + threadIdx.x;
    bool valid = e < n;
    uint32_t x = valid ? v[e] : 0u;
    uint64_t keep = (valid && (e == 0 || v[e - 1] != x) && !is_removed(a.rem, x)) ? 1u : 0u;
    uint64_t tot;
    uint64_t ex = block_exclusive_scan(keep, s_ws, tot);  // syncs: reads precede writes
    if (keep) v[outn + ex] = x;
    outn += tot;
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  if (warp_id() == 0) {
    uint32_t enc = a.want_enc ? intcomp::enc_size_warp(v, (uint32_t)outn) : 0u;
    if (lane_id() == 0) s_enc = enc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.g_cnt[hp] = (uint32_t)outn;
    a.g_enc[hp] = s_enc;
    if (outn || a.keep_empty) {
      uint32_t lo = 0, hi = a.nb1 - 1;  // bucket of hp: last b with bk_pos[b] <= hp
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a.bk_pos[mid + 1] <= hp)
          lo = mid + 1;
        else
          hi = mid;
      }
      const uint32_t b = lo;
      unsigned long long* bo = reinterpret_cast<unsigned long long*>(a.bk_out);
      atomicAdd(&bo[0ull * a.nb1 + b], 1ull);
      atomicAdd(&bo[1ull * a.nb1 + b],
                (unsigned long long)term_len_of(a.segs, a.k, a.ord_inst[hp]));
      atomicAdd(&bo[2ull * a.nb1 + b], (unsigned long long)outn);
      atomicAdd(&bo[3ull * a.nb1 + b], (unsigned long long)s_enc);
    }
  }
}

// ---------------------------------------------------------------- host driver
int k2_union(const MergePlan& plan, const RemovedSet& rem, bool want_enc, bool keep_empty,
             uint64_t n_in_hint, UnionOut& u, cudaStream_t s) {
  u.keep_empty = keep_empty;
  const uint32_t B = plan.n_buckets, N = plan.n_total;
  uint64_t n_in = n_in_hint;
  if (n_in == 0) {  // Σ input postings inside the windows, to size tmp_post
    uint64_t h_plan_tot[2] = {0, 0};
    II2_CUDA_TRY(cudaMemcpyAsync(h_plan_tot, plan.totals.p, 16, cudaMemcpyDeviceToHost, s));
    II2_CUDA_TRY(cudaStreamSynchronize(s));
    n_in = h_plan_tot[0];
  }
  II2_TRY(u.tmp_post.alloc(n_in, s));
  II2_TRY(u.g_cnt.alloc(N, s));
  II2_TRY(u.g_enc.alloc(N, s));
  II2_TRY(u.g_off.alloc(N, s));
  II2_TRY(u.bk_raw.alloc(4 * (size_t)(B + 1), s));
  II2_TRY(u.bk_out.alloc(4 * (size_t)(B + 1), s));
  II2_TRY(u.totals.alloc(5, s));  // [4] totals + n_large
  DevBuf<uint32_t> large_pos;
  DevBuf<uint64_t> large_len;
  const uint32_t large_cap = (uint32_t)(n_in / K2_WCAP + 1);  // each deferred group has L > WCAP
  II2_TRY(large_pos.alloc(large_cap, s));
  II2_TRY(large_len.alloc(large_cap, s));
  uint32_t* d_n_large = reinterpret_cast<uint32_t*>(u.totals.p + 4);
  II2_CUDA_TRY(cudaMemsetAsync(u.totals.p + 4, 0, 8, s));
  II2_CUDA_TRY(cudaMemsetAsync(u.bk_raw.p, 0, 4 * (size_t)(B + 1) * 8, s));

  K2Args a;
  a.segs = plan.segs;
  a.k = plan.k;
  a.bk_pos = plan.bk_pos.p;
  a.bk_Pbase = plan.bk_P();
  a.ord_inst = plan.ord_inst.p;
  a.src_ptr = plan.src_ptr.p;
  a.src_len = plan.src_len.p;
  a.gsz = plan.gsz.p;
  a.rem = rem;
  a.want_enc = want_enc ? 1 : 0;
  a.keep_empty = keep_empty ? 1 : 0;
  a.tmp_post = u.tmp_post.p;
  a.g_cnt = u.g_cnt.p;
  a.g_enc = u.g_enc.p;
  a.g_off = u.g_off.p;
  a.bk_out = u.bk_raw.p;
  a.nb1 = B + 1;
  a.large_pos = large_pos.p;
  a.large_len = large_len.p;
  a.n_large = d_n_large;

  {
    ProfScope scope("k2_union", s);
    k2_union_kernel<<<B, K2_THREADS, 0, s>>>(a);
    II2_LAUNCHED();
  }
  // optimistic: scan right away; redone only if large groups were deferred
  II2_TRY(exclusive_scan_multi_u64(u.bk_raw.p, u.bk_out.p, B + 1, 4, u.totals.p, s));
  uint64_t h_tot[5];
  II2_CUDA_TRY(cudaMemcpyAsync(h_tot, u.totals.p, 40, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  const uint32_t h_nl = (uint32_t)h_tot[4];
  if (h_nl > 0) {
    ProfScope scope("k2_large", s);
    std::vector<uint64_t> lens(h_nl);
    II2_CUDA_TRY(cudaMemcpyAsync(lens.data(), large_len.p, (size_t)h_nl * 8,
                                 cudaMemcpyDeviceToHost, s));
    II2_CUDA_TRY(cudaStreamSynchronize(s));
    for (uint64_t x : lens) {
      if (x >= (1ull << 32)) {
        set_last_error("a single term unions %llu postings (max 2^32-1)", (unsigned long long)x);
        return II2_ERR_UNSUPPORTED;
      }
    }
    for (uint32_t y0 = 0; y0 < h_nl; y0 += 32768) {  // grid.y limit
      const uint32_t ny = std::min<uint32_t>(32768, h_nl - y0);
      uint64_t maxL = 0;
      for (uint32_t i = 0; i < ny; i++) maxL = std::max(maxL, lens[y0 + i]);
      K2Args b2 = a;
      b2.large_pos = a.large_pos + y0;
      b2.large_len = a.large_len + y0;
      const unsigned gx = (unsigned)std::min<uint64_t>((maxL + 4095) / 4096, 2048);
      dim3 grid(gx, ny);
      k2_large_gather<<<grid, 256, 0, s>>>(b2);
      II2_LAUNCHED();
      k2_large_tile_sort<<<grid, 512, 0, s>>>(b2);
      II2_LAUNCHED();
      for (uint64_t kk = 2ull * LG_TILE; (kk >> 1) < maxL; kk <<= 1) {
        k2_large_stage<<<grid, 256, 0, s>>>(b2, kk, 0);
        II2_LAUNCHED();
        for (uint64_t j = kk >> 2; j >= LG_TILE; j >>= 1) {
          k2_large_stage<<<grid, 256, 0, s>>>(b2, kk, j);
          II2_LAUNCHED();
        }
        k2_large_tile_merge<<<grid, 512, 0, s>>>(b2, kk);
        II2_LAUNCHED();
      }
      k2_large_finish<<<ny, 1024, 0, s>>>(b2);
      II2_LAUNCHED();
    }
    II2_TRY(exclusive_scan_multi_u64(u.bk_raw.p, u.bk_out.p, B + 1, 4, u.totals.p, s));
    II2_CUDA_TRY(cudaMemcpyAsync(h_tot, u.totals.p, 32, cudaMemcpyDeviceToHost, s));
    II2_CUDA_TRY(cudaStreamSynchronize(s));
  }
  for (int i = 0; i < 4; i++) u.h_totals[i] = h_tot[i];
  return II2_OK;
}

}  // namespace ii2
