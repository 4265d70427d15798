// k12_union.cu — K12: the k-way term merge finished per bucket, fused with the per-term union +
// dedup of uint32 posting lists, the removed filter and the intcomp encoder.
//
// Replaces, in one pass over the inputs:
//   - go-iterators' MergingIterator (built at shard.go:267; ordering file.CompareTermValues,
//     file/types.go:24-26): inside a bucket, equal terms are grouped with a shared-memory hash
//     table and only the DISTINCT terms are sorted (16-byte key windows past the bucket's common
//     prefix, bitonic network);
//   - file.MergeTermValues (file/types.go:14-22: append + slices.Sort + slices.Compact, applied
//     pairwise) and the removed filter of the merge loop (shard.go:181-190);
//   - the size half and the byte half of intcomp.CompressUint32 (file/writer.go:49).
// Semantics kept bit-exact:
//   - a term present in ONE segment passes through untouched — not sorted, not deduped (survey
//     Q4); a term present in >= 2 segments becomes the sorted-unique union;
//   - the filter runs AFTER the union; removed = membership in the sorted removed list.
//
// One CTA per bucket (k1_plan.cu).  A bucket is consumed in sub-tiles of <= 1024 instances
// (oversized buckets are bisected on the fly, pivot = median of the widest run).  Per sub-tile:
// keys -> hash grouping -> sort of the distinct terms -> instances laid out group by group ->
// postings gathered from HBM into shared memory in grouped order (batches of <= 4096 values)
// -> one warp per term: bitonic sort in registers (<= 256 values), dedup against the neighbour
// lane, removed filter (L2-resident bitmap), ballot compaction -> encoded size -> one
// atomicAdd per batch reserves the `_val` words -> encode from shared memory.
// Terms whose lists exceed 4096 values go to the multi-CTA global-memory path at the end of
// this file.  Integer/byte work, HBM-bound by design: every input byte is read once.
#include <algorithm>
#include <vector>

#include "intcomp.cuh"
#include "keys.cuh"
#include "union.cuh"

namespace ii2 {

constexpr int K12_THREADS = 256;
constexpr int K12_WARPS = K12_THREADS / 32;
constexpr uint32_t CAP_I = 1024;   // instances per sub-tile
constexpr uint32_t CAP_P = 4096;   // postings per batch (shared-memory union buffer)
constexpr uint32_t K12_HT = 2048;  // hash slots
constexpr uint32_t REG_CAP = 256;  // values a warp sorts in registers
constexpr uint16_t K12_EMPTY = 0xFFFFu;
constexpr uint32_t K12_PENDING = 0xFFFFFFFFu;

struct K12Args {
  const SegDesc* segs;
  int k;
  const uint32_t* part;
  const uint32_t* row_of;
  const uint64_t* bk_pos;
  const uint64_t* bk_P;
  const uint32_t* bk_cpl;
  RemovedSet rem;
  int want_enc, want_dec, keep_empty;
  GroupRec* recs;
  uint32_t* tmp_post;
  uint32_t* tmp_enc;
  unsigned long long* enc_alloc;
  uint64_t* bk_raw;  // [4][nb1]
  uint32_t* bk_D;
  uint32_t nb1;
  // heavy terms, deferred to the multi-CTA path
  uint32_t* n_large;
  uint32_t* large_rec;
  uint32_t* large_beg;
  uint32_t* large_c;
  uint32_t* large_bucket;
  unsigned long long* n_lsrc;
  uint64_t* lsrc_ptr;
  uint64_t* lsrc_len;
};

__host__ __device__ inline size_t k12_smem_bytes(int k) {
  return (size_t)CAP_I * 16                      // key_hi, key_lo
         + (size_t)CAP_I * 4 * 3 + (size_t)(CAP_I + 1) * 4   // idx, plen, gl, cnt/ppre
         + (size_t)CAP_P * 4                     // res
         + (size_t)(5 * k + 1) * 4               // cur, mm, hi, endr, rstart
         + (size_t)K12_WARPS * intcomp::kStageWords * 4
         + (size_t)CAP_I * 2 * 6 + (size_t)(CAP_I + 1) * 2    // seg, tlen, grp/queue, gstart, reps, order, gss
         + (size_t)K12_HT * 2 + 64;
}

// ---- warp-level union of one group held in shared memory --------------------------------
// Bitonic sort of 32*R values striped over the warp (element e = r*32 + lane), then dedup
// (slices.Compact) + removed filter, compacted back to base[0..outn).  Returns outn.
template <int R>
__device__ __forceinline__ uint32_t union_regs(uint32_t* base, uint32_t L, const RemovedSet& rem) {
  const unsigned lane = lane_id();
  uint32_t v[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t e = r * 32 + lane;
    v[r] = e < L ? base[e] : 0xFFFFFFFFu;  // padding sorts to the end; only L values are used
  }
#pragma unroll
  for (uint32_t kk = 2; kk <= 32u * R; kk <<= 1) {
#pragma unroll
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; r++) {
          if ((r & dr) == 0) {
            const uint32_t e = r * 32 + lane;
            const bool asc = (e & kk) == 0;
            const uint32_t x = v[r], y = v[r | dr];
            const bool sw = asc ? (y < x) : (x < y);
            v[r] = sw ? y : x;
            v[r | dr] = sw ? x : y;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
          const uint32_t e = r * 32 + lane;
          const bool asc = (e & kk) == 0;
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool lower = (lane & j) == 0;
          const uint32_t mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
          v[r] = (lower == asc) ? mn : mx;
        }
      }
    }
  }
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  uint32_t outn = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t e = r * 32 + lane;
    uint32_t prev = __shfl_up_sync(0xffffffffu, v[r], 1);
    if (r > 0) {
      const uint32_t p31 = __shfl_sync(0xffffffffu, v[r - 1], 31);
      if (lane == 0) prev = p31;
    }
    const bool keep = e < L && (e == 0 || prev != v[r]) && !is_removed(rem, v[r]);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) base[outn + __popc(bal & lt)] = v[r];
    outn += __popc(bal);
  }
  __syncwarp();
  return outn;
}

// single-source pass-through: order and duplicates kept, only the filter (in place)
__device__ __forceinline__ uint32_t filter_inplace_warp(uint32_t* base, uint32_t L,
                                                        const RemovedSet& rem) {
  if (rem.n == 0) return L;
  const unsigned lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  uint32_t outn = 0;
  for (uint32_t e0 = 0; e0 < L; e0 += 32) {
    const uint32_t e = e0 + lane;
    const bool valid = e < L;
    const uint32_t v = valid ? base[e] : 0u;
    const bool keep = valid && !is_removed(rem, v);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) base[outn + __popc(bal & lt)] = v;
    outn += __popc(bal);
    __syncwarp();
  }
  return outn;
}

__global__ void __launch_bounds__(K12_THREADS, 3) k12_kernel(const K12Args a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint64_t s_ws64[K12_WARPS + 2];
  __shared__ uint32_t s_ws32[K12_WARPS + 2];
  __shared__ unsigned long long s_acc[4];
  __shared__ unsigned long long s_alloc;
  __shared__ uint32_t s_nq, s_ncta;

  const int k = a.k;
  uint8_t* sp = smem_raw;
  uint64_t* key_hi = reinterpret_cast<uint64_t*>(sp); sp += CAP_I * 8;
  uint64_t* key_lo = reinterpret_cast<uint64_t*>(sp); sp += CAP_I * 8;
  uint32_t* idx_a = reinterpret_cast<uint32_t*>(sp); sp += CAP_I * 4;
  uint32_t* plen = reinterpret_cast<uint32_t*>(sp); sp += CAP_I * 4;
  uint32_t* gl = reinterpret_cast<uint32_t*>(sp); sp += CAP_I * 4;      // by representative
  uint32_t* cnt = reinterpret_cast<uint32_t*>(sp); sp += (CAP_I + 1) * 4;  // by representative
  uint32_t* res = reinterpret_cast<uint32_t*>(sp); sp += CAP_P * 4;
  uint32_t* cur = reinterpret_cast<uint32_t*>(sp); sp += k * 4;
  uint32_t* mm = reinterpret_cast<uint32_t*>(sp); sp += k * 4;
  uint32_t* hib = reinterpret_cast<uint32_t*>(sp); sp += k * 4;
  uint32_t* endr = reinterpret_cast<uint32_t*>(sp); sp += k * 4;
  uint32_t* rstart = reinterpret_cast<uint32_t*>(sp); sp += (k + 1) * 4;
  uint32_t* stage = reinterpret_cast<uint32_t*>(sp); sp += K12_WARPS * intcomp::kStageWords * 4;
  uint16_t* seg_a = reinterpret_cast<uint16_t*>(sp); sp += CAP_I * 2;
  uint16_t* tlen = reinterpret_cast<uint16_t*>(sp); sp += CAP_I * 2;
  uint16_t* grp = reinterpret_cast<uint16_t*>(sp); sp += CAP_I * 2;
  uint16_t* gstart = reinterpret_cast<uint16_t*>(sp); sp += CAP_I * 2;  // by representative
  uint16_t* reps = reinterpret_cast<uint16_t*>(sp); sp += CAP_I * 2;    // by rank
  uint16_t* order = reinterpret_cast<uint16_t*>(sp); sp += CAP_I * 2;   // by position
  uint16_t* gss = reinterpret_cast<uint16_t*>(sp); sp += (CAP_I + 1) * 2 + 2;  // by rank
  uint16_t* table = reinterpret_cast<uint16_t*>(sp);
  // aliases, valid once the phase that owns the original is over
  uint32_t* ppre = cnt;                                   // by position, after the scatter
  uint16_t* queue = grp;                                  // after the position scan
  uint32_t* gout = reinterpret_cast<uint32_t*>(key_hi);   // by rank, after the sort
  uint32_t* genc = gout + CAP_I;
  uint32_t* exenc = reinterpret_cast<uint32_t*>(key_lo);
  uint16_t* ctaq = reinterpret_cast<uint16_t*>(exenc + CAP_I);

  const uint32_t tid = threadIdx.x;
  const unsigned lane = lane_id(), warp = warp_id();
  const uint32_t b = blockIdx.x;
  uint32_t W = (uint32_t)(a.bk_pos[b + 1] - a.bk_pos[b]);
  if (W == 0) {
    if (tid < 4) a.bk_raw[(uint64_t)tid * a.nb1 + b] = 0;
    if (tid == 0) a.bk_D[b] = 0;
    return;
  }
  {
    const uint32_t r0 = a.row_of[b], r1 = a.row_of[b + 1];
    for (int s = tid; s < k; s += K12_THREADS) {
      cur[s] = a.part[(uint64_t)r0 * k + s];
      endr[s] = a.part[(uint64_t)r1 * k + s];
    }
  }
  if (tid < 4) s_acc[tid] = 0;
  __syncthreads();
  const uint64_t rec_base = a.bk_pos[b];
  const uint64_t Pbase = a.bk_P[b];
  const uint32_t cpl = a.bk_cpl[b];
  uint32_t dcount = 0;       // distinct terms emitted so far in this bucket
  uint64_t tile_in_base = 0;  // input postings of the previous sub-tiles (light groups)

  while (W > 0) {
    // ---------------- choose the sub-tile [cur, mm) ----------------
    uint32_t size;
    if (W <= CAP_I) {
      for (int s = tid; s < k; s += K12_THREADS) mm[s] = endr[s];
      size = W;
      __syncthreads();
    } else {
      for (int s = tid; s < k; s += K12_THREADS) hib[s] = endr[s];
      __syncthreads();
      for (;;) {
        uint64_t best = 0;  // widest run and its median term = pivot
        for (int s = tid; s < k; s += K12_THREADS) {
          const uint64_t cand = ((uint64_t)(hib[s] - cur[s]) << 32) | (uint32_t)s;
          best = cand > best ? cand : best;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const uint64_t o = __shfl_xor_sync(0xffffffffu, best, d);
          best = o > best ? o : best;
        }
        __syncthreads();
        if (lane == 0) s_ws64[warp] = best;
        __syncthreads();
        for (int w2 = 0; w2 < K12_WARPS; w2++) best = s_ws64[w2] > best ? s_ws64[w2] : best;
        const int spv = (int)(uint32_t)best;
        const uint32_t win = (uint32_t)(best >> 32);  // >= 2: sum of windows > CAP_I >= k
        const uint32_t mid = cur[spv] + (win >> 1);
        const KeyedTerm pivot = keyed_term(a.segs[spv], mid);
        uint32_t part_sum = 0;
        for (int s = tid; s < k; s += K12_THREADS) {
          const uint32_t m = s == spv ? mid : keyed_lower_bound(a.segs[s], cur[s], hib[s], pivot);
          mm[s] = m;
          part_sum += m - cur[s];
        }
        uint32_t tot;
        block_exclusive_scan(part_sum, s_ws32, tot);
        size = tot;  // >= 1: the pivot's own run contributes mid - cur >= 1
        if (size <= CAP_I) break;
        for (int s = tid; s < k; s += K12_THREADS) hib[s] = mm[s];
        __syncthreads();
      }
      __syncthreads();
    }

    // ---------------- (1) run starts ----------------
    {
      uint32_t run = 0;
      for (int base = 0; base < k; base += K12_THREADS) {
        const int s = base + tid;
        const uint32_t v = s < k ? mm[s] - cur[s] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan(v, s_ws32, tot);
        if (s < k) rstart[s] = run + ex;
        run += tot;
      }
      if (tid == 0) rstart[k] = size;
    }
    for (uint32_t t = tid; t < K12_HT; t += K12_THREADS) table[t] = K12_EMPTY;
    __syncthreads();

    // ---------------- (2) key windows + posting lengths ----------------
    for (uint32_t i = tid; i < size; i += K12_THREADS) {
      int lo = 0, hi = k;  // first s with rstart[s+1] > i
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rstart[mid + 1] <= i)
          lo = mid + 1;
        else
          hi = mid;
      }
      const int s = lo;
      const uint32_t idx = cur[s] + (i - rstart[s]);
      const SegDesc& sd = a.segs[s];
      const uint32_t o = __ldg(sd.toff + idx), n = __ldg(sd.toff + idx + 1) - o;
      const uint64_t p0 = __ldg(sd.poff + idx), p1 = __ldg(sd.poff + idx + 1);
      uint64_t kh, kl;
      load_key16(sd.tb, o, n, cpl, kh, kl);
      key_hi[i] = kh;
      key_lo[i] = kl;
      tlen[i] = (uint16_t)n;
      idx_a[i] = idx;
      seg_a[i] = (uint16_t)s;
      plen[i] = (p1 - p0) > 0xFFFFFFFEull ? 0xFFFFFFFFu : (uint32_t)(p1 - p0);
      cnt[i] = 0;
      gl[i] = 0;
    }
    __syncthreads();

    // bytes past the 16-byte window, only needed for terms longer than cpl+16
    auto tail_compare = [&](uint32_t x, uint32_t y) -> int {
      const uint32_t skip = cpl + 16;
      const uint32_t nx = tlen[x], ny = tlen[y];
      if (nx > skip && ny > skip) {
        const SegDesc& sx = a.segs[seg_a[x]];
        const SegDesc& sy = a.segs[seg_a[y]];
        const uint8_t* px = sx.tb + __ldg(sx.toff + idx_a[x]) + skip;
        const uint8_t* py = sy.tb + __ldg(sy.toff + idx_a[y]) + skip;
        return term_compare(px, nx - skip, py, ny - skip);
      }
      return nx < ny ? -1 : (nx > ny ? 1 : 0);
    };

    // ---------------- (3) group equal terms (hash table of representatives) ----------------
    for (uint32_t i = tid; i < size; i += K12_THREADS) {
      const uint64_t kh = key_hi[i], kl = key_lo[i];
      uint64_t h = kh * 0x9E3779B97F4A7C15ull;
      h ^= (kl + 0xD6E8FEB86659FD93ull + (h << 6) + (h >> 2));
      h *= 0xFF51AFD7ED558CCDull;
      h ^= h >> 33;
      h += tlen[i] * 0xC2B2AE3D27D4EB4Full;
      h ^= h >> 29;
      uint32_t slot = (uint32_t)h & (K12_HT - 1);
      uint32_t rep;
      for (;;) {
        const unsigned short prev = atomicCAS(reinterpret_cast<unsigned short*>(&table[slot]),
                                              (unsigned short)K12_EMPTY, (unsigned short)i);
        if (prev == K12_EMPTY) {
          rep = i;
          break;
        }
        if (key_hi[prev] == kh && key_lo[prev] == kl && tlen[prev] == tlen[i] &&
            tail_compare(i, prev) == 0) {
          rep = prev;
          break;
        }
        slot = (slot + 1) & (K12_HT - 1);
      }
      grp[i] = (uint16_t)rep;
      atomicAdd(&cnt[rep], 1u);
      // group length, saturating per source so the sum cannot wrap: heavy iff sum > CAP_P
      atomicAdd(&gl[rep], plen[i] > CAP_P ? CAP_P + 1 : plen[i]);
    }
    __syncthreads();

    // ---------------- (4) list of distinct representatives ----------------
    uint32_t D = 0;
    for (uint32_t base = 0; base < size; base += K12_THREADS) {
      const uint32_t i = base + tid;
      const uint32_t f = (i < size && grp[i] == i) ? 1u : 0u;
      uint32_t tot;
      const uint32_t ex = block_exclusive_scan(f, s_ws32, tot);
      if (f) reps[D + ex] = (uint16_t)i;
      D += tot;
    }
    __syncthreads();

    // ---------------- (5) sort the distinct terms ----------------
    {
      auto less = [&](uint16_t x, uint16_t y) -> bool {
        const uint64_t hx = key_hi[x], hy = key_hi[y];
        if (hx != hy) return hx < hy;
        const uint64_t lx = key_lo[x], ly = key_lo[y];
        if (lx != ly) return lx < ly;
        return tail_compare(x, y) < 0;
      };
      if (D <= 32) {
        if (warp == 0) bitonic_sort_any(reps, D, lane, 32u, less, [] { __syncwarp(); });
      } else {
        bitonic_sort_any(reps, D, tid, (uint32_t)K12_THREADS, less, [] { __syncthreads(); });
      }
    }
    __syncthreads();

    // ---------------- (6) group start positions, in sorted order ----------------
    {
      uint32_t run = 0;
      for (uint32_t base = 0; base < D; base += K12_THREADS) {
        const uint32_t r = base + tid;
        const uint32_t c = r < D ? cnt[reps[r]] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan(c, s_ws32, tot);
        if (r < D) {
          gstart[reps[r]] = (uint16_t)(run + ex);
          gss[r] = (uint16_t)(run + ex);
        }
        run += tot;
      }
      if (tid == 0) gss[D] = (uint16_t)size;
    }
    __syncthreads();

    // ---------------- (7) lay the instances out group by group ----------------
    for (uint32_t i = tid; i < size; i += K12_THREADS) {
      const uint32_t g = grp[i];
      const uint32_t slot = atomicSub(&cnt[g], 1u) - 1u;
      order[gstart[g] + slot] = (uint16_t)i;
    }
    __syncthreads();

    // ---------------- (8) posting offsets in grouped order (heavy groups take no room) -------
    {
      constexpr uint32_t PER = CAP_I / K12_THREADS;
      uint32_t e[PER];
      uint32_t sum = 0;
#pragma unroll
      for (uint32_t j = 0; j < PER; j++) {
        const uint32_t p = tid * PER + j;
        e[j] = 0;
        if (p < size) {
          const uint32_t i = order[p];
          e[j] = gl[grp[i]] > CAP_P ? 0u : plen[i];
        }
        sum += e[j];
      }
      uint32_t tot;
      uint32_t ex = block_exclusive_scan(sum, s_ws32, tot);  // syncs: cnt is dead, ppre takes over
#pragma unroll
      for (uint32_t j = 0; j < PER; j++) {
        const uint32_t p = tid * PER + j;
        if (p < size) ppre[p] = ex;
        ex += e[j];
      }
      if (tid == 0) ppre[size] = tot;
    }
    __syncthreads();
    const uint32_t tile_in = ppre[size];

    // ---------------- (9) batches of groups whose postings fit the union buffer -------------
    uint32_t r0 = 0;
    while (r0 < D) {
      const uint32_t pbase = ppre[gss[r0]];
      uint32_t r1;
      {
        uint32_t ok = 0;
        for (uint32_t r = r0 + 1 + tid; r <= D; r += K12_THREADS)
          if (ppre[gss[r]] - pbase <= CAP_P) ok++;
        uint32_t tot;
        block_exclusive_scan(ok, s_ws32, tot);
        r1 = r0 + tot;  // >= r0 + 1: a light group alone always fits
      }
      if (tid == 0) {
        s_nq = 0;
        s_ncta = 0;
      }
      __syncthreads();
      // ---- gather: postings of every instance to its place in grouped order ----
      const uint32_t pend = gss[r1];
      for (uint32_t p = gss[r0] + tid; p < pend; p += K12_THREADS) {
        const uint32_t n = ppre[p + 1] - ppre[p];
        if (n == 0) continue;
        const uint32_t i = order[p];
        const SegDesc& sd = a.segs[seg_a[i]];
        const uint32_t* src = sd.post + __ldg(sd.poff + idx_a[i]);
        uint32_t* dst = res + (ppre[p] - pbase);
        const uint32_t m = n < 32u ? n : 32u;
        for (uint32_t t = 0; t < m; t++) dst[t] = __ldg(src + t);
        if (n > 32u) queue[atomicAdd(&s_nq, 1u)] = (uint16_t)p;
      }
      __syncthreads();
      for (uint32_t q = warp; q < s_nq; q += K12_WARPS) {  // long lists: one warp each
        const uint32_t p = queue[q];
        const uint32_t n = ppre[p + 1] - ppre[p];
        const uint32_t i = order[p];
        const SegDesc& sd = a.segs[seg_a[i]];
        const uint32_t* src = sd.post + __ldg(sd.poff + idx_a[i]);
        uint32_t* dst = res + (ppre[p] - pbase);
        for (uint32_t t = 32 + lane; t < n; t += 32) dst[t] = __ldg(src + t);
      }
      __syncthreads();

      // ---- union: one warp per term ----
      for (uint32_t r = r0 + warp; r < r1; r += K12_WARPS) {
        const uint32_t ps = gss[r], c = gss[r + 1] - ps;
        const uint32_t L = ppre[gss[r + 1]] - ppre[ps];
        uint32_t* base = res + (ppre[ps] - pbase);
        if (gl[reps[r]] > CAP_P) {  // heavy: hand the sources to the multi-CTA path
          uint32_t slot = 0, beg = 0;
          if (lane == 0) {
            slot = atomicAdd(a.n_large, 1u);
            beg = (uint32_t)atomicAdd(a.n_lsrc, (unsigned long long)c);
            a.large_rec[slot] = (uint32_t)(rec_base + dcount + r);
            a.large_beg[slot] = beg;
            a.large_c[slot] = c;
            a.large_bucket[slot] = b;
            gout[r] = K12_PENDING;
            genc[r] = 0;
          }
          beg = __shfl_sync(0xffffffffu, beg, 0);
          for (uint32_t j = lane; j < c; j += 32) {
            const uint32_t i = order[ps + j];
            const SegDesc& sd = a.segs[seg_a[i]];
            const uint64_t p0 = __ldg(sd.poff + idx_a[i]), p1 = __ldg(sd.poff + idx_a[i] + 1);
            a.lsrc_ptr[beg + j] = reinterpret_cast<uint64_t>(sd.post + p0);
            a.lsrc_len[beg + j] = p1 - p0;
          }
          continue;
        }
        uint32_t outn;
        if (c == 1) {
          outn = filter_inplace_warp(base, L, a.rem);
        } else if (L <= 32) {
          outn = union_regs<1>(base, L, a.rem);
        } else if (L <= 64) {
          outn = union_regs<2>(base, L, a.rem);
        } else if (L <= 128) {
          outn = union_regs<4>(base, L, a.rem);
        } else if (L <= REG_CAP) {
          outn = union_regs<8>(base, L, a.rem);
        } else {
          if (lane == 0) ctaq[atomicAdd(&s_ncta, 1u)] = (uint16_t)r;
          continue;
        }
        const uint32_t enc = a.want_enc ? intcomp::enc_size_warp(base, outn) : 0u;
        if (lane == 0) {
          gout[r] = outn;
          genc[r] = enc;
        }
      }
      __syncthreads();
      // ---- terms too long for a warp's registers: the whole CTA, in place ----
      for (uint32_t q = 0; q < s_ncta; q++) {
        const uint32_t r = ctaq[q];
        const uint32_t ps = gss[r];
        const uint32_t L = ppre[gss[r + 1]] - ppre[ps];
        uint32_t* base = res + (ppre[ps] - pbase);
        bitonic_sort_any(base, L, tid, (uint32_t)K12_THREADS,
                         [](uint32_t x, uint32_t y) { return x < y; }, [] { __syncthreads(); });
        __syncthreads();
        uint32_t outn = 0;
        for (uint32_t e0 = 0; e0 < L; e0 += K12_THREADS) {
          const uint32_t e = e0 + tid;
          const bool valid = e < L;
          const uint32_t v = valid ? base[e] : 0u;
          const uint32_t keep =
              (valid && (e == 0 || base[e - 1] != v) && !is_removed(a.rem, v)) ? 1u : 0u;
          uint32_t tot;
          const uint32_t ex = block_exclusive_scan(keep, s_ws32, tot);  // reads precede writes
          if (keep) base[outn + ex] = v;
          outn += tot;
          __syncthreads();
        }
        if (warp == 0) {
          const uint32_t enc = a.want_enc ? intcomp::enc_size_warp(base, outn) : 0u;
          if (lane == 0) {
            gout[r] = outn;
            genc[r] = enc;
          }
        }
        __syncthreads();
      }

      // ---- totals of the batch, `_val` words reserved with one atomicAdd ----
      {
        uint64_t run_e = 0;
        uint64_t tot_pack = 0;
        for (uint32_t base = r0; base < r1; base += K12_THREADS) {
          const uint32_t r = base + tid;
          uint64_t pack = 0;
          uint64_t e = 0;
          if (r < r1) {
            const uint32_t o = gout[r];
            if (o != K12_PENDING && (o || a.keep_empty)) {
              pack = 1ull | ((uint64_t)tlen[reps[r]] << 12) | ((uint64_t)o << 40);
              e = genc[r];
            }
          }
          uint64_t tp, te;
          block_exclusive_scan(pack, s_ws64, tp);
          const uint64_t ex = block_exclusive_scan(e, s_ws64, te);
          if (r < r1) exenc[r] = (uint32_t)(run_e + ex);
          run_e += te;
          tot_pack += tp;
        }
        if (tid == 0) {
          s_acc[0] += tot_pack & 0xFFFull;
          s_acc[1] += (tot_pack >> 12) & 0xFFFFFFFull;
          s_acc[2] += tot_pack >> 40;
          s_acc[3] += run_e;
          s_alloc = (a.want_enc && run_e) ? atomicAdd(a.enc_alloc, (unsigned long long)run_e) : 0ull;
        }
      }
      __syncthreads();

      // ---- emit: record per term, decoded copy and/or encoded stream ----
      for (uint32_t r = r0 + warp; r < r1; r += K12_WARPS) {
        const uint32_t ps = gss[r];
        const uint32_t outn = gout[r];
        const uint32_t* base = res + (ppre[ps] - pbase);
        uint32_t* dec = a.want_dec ? a.tmp_post + Pbase + tile_in_base + ppre[ps] : nullptr;
        const uint64_t eoff = s_alloc + exenc[r];
        if (lane == 0) {
          const uint32_t i = reps[r];
          const SegDesc& sd = a.segs[seg_a[i]];
          GroupRec g;
          g.dec = reinterpret_cast<uint64_t>(dec);
          g.eoff = eoff;
          g.inst = sd.base + (idx_a[i] - sd.lo);
          g.tlen = tlen[i];
          g.cnt = outn;
          g.enc = genc[r];
          a.recs[rec_base + dcount + r] = g;
        }
        if (outn == K12_PENDING || outn == 0) continue;
        if (a.want_dec)
          for (uint32_t e = lane; e < outn; e += 32) dec[e] = base[e];
        if (a.want_enc)
          intcomp::enc_emit_warp(base, outn, a.tmp_enc + eoff, stage + warp * intcomp::kStageWords);
      }
      __syncthreads();
      r0 = r1;
    }

    dcount += D;
    tile_in_base += tile_in;
    W -= size;
    for (int s = tid; s < k; s += K12_THREADS) cur[s] = mm[s];
    __syncthreads();
  }
  if (tid < 4) a.bk_raw[(uint64_t)tid * a.nb1 + b] = s_acc[tid];
  if (tid == 0) a.bk_D[b] = dcount;
}

// ---------------------------------------------------------------- heavy terms (global memory)
struct LargeArgs {
  const uint32_t* rec;     // record index of the term
  const uint32_t* beg;     // first source in lsrc_*
  const uint32_t* c;       // number of sources
  const uint32_t* bucket;
  const uint64_t* lsrc_ptr;
  const uint64_t* lsrc_len;
  uint64_t* len;           // Σ source lengths
  const uint64_t* off;     // offset into tmp
  uint32_t* tmp;
  GroupRec* recs;
  RemovedSet rem;
  int want_enc, keep_empty;
  uint32_t* tmp_enc;
  unsigned long long* enc_alloc;
  uint64_t* bk_raw;
  uint32_t nb1;
};

__global__ void __launch_bounds__(256) k2_large_len(const LargeArgs a, uint32_t n) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= n) return;
  uint64_t L = 0;
  for (uint32_t j = lane_id(); j < a.c[g]; j += 32) L += a.lsrc_len[a.beg[g] + j];
  L = warp_sum(L);
  if (lane_id() == 0) a.len[g] = L;
}

// gather: grid (x = CTAs per group, y = group)
__global__ void __launch_bounds__(256) k2_large_gather(const LargeArgs a) {
  __shared__ uint64_t s_moff[kMaxSegs + 1];
  __shared__ uint64_t s_ws[256 / 32 + 2];
  const uint32_t g = blockIdx.y;
  const uint32_t c = a.c[g], beg = a.beg[g];
  uint64_t run = 0;
  for (uint32_t base = 0; base < c; base += 256) {
    const uint32_t i = base + threadIdx.x;
    const uint64_t li = i < c ? a.lsrc_len[beg + i] : 0u;
    uint64_t tot;
    const uint64_t ex = block_exclusive_scan(li, s_ws, tot);
    if (i < c) s_moff[i] = run + ex;
    run += tot;
  }
  if (threadIdx.x == 0) s_moff[c] = run;
  __syncthreads();
  const uint64_t n = run;
  uint32_t* dst = a.tmp + a.off[g];
  for (uint64_t e = (uint64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (uint64_t)gridDim.x * 256) {
    uint32_t lo = 0, hi = c;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_moff[mid + 1] <= e)
        lo = mid + 1;
      else
        hi = mid;
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.lsrc_ptr[beg + lo]);
    dst[e] = __ldg(src + (e - s_moff[lo]));
  }
}

constexpr uint32_t LG_TILE = 4096;

// sort every aligned LG_TILE tile of every heavy group in shared memory
__global__ void __launch_bounds__(512) k2_large_tile_sort(const LargeArgs a) {
  __shared__ uint32_t tile[LG_TILE];
  if (a.c[blockIdx.y] == 1) return;  // single source: passes through unsorted (survey Q4)
  const uint64_t n = a.len[blockIdx.y];
  uint32_t* base = a.tmp + a.off[blockIdx.y];
  for (uint64_t t0 = (uint64_t)blockIdx.x * LG_TILE; t0 < n; t0 += (uint64_t)gridDim.x * LG_TILE) {
    const uint32_t m = (uint32_t)((n - t0) < LG_TILE ? (n - t0) : LG_TILE);
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
__global__ void __launch_bounds__(256) k2_large_stage(const LargeArgs a, uint64_t kk, uint64_t j) {
  const uint64_t n = a.len[blockIdx.y];
  if ((kk >> 1) >= n || a.c[blockIdx.y] == 1) return;  // sorted at this block size / pass-through
  uint32_t* v = a.tmp + a.off[blockIdx.y];
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
      const uint32_t x = v[i], y = v[l];
      if (y < x) {
        v[i] = y;
        v[l] = x;
      }
    }
  }
}

// finish block size kk inside shared memory: half-cleaners j = LG_TILE/2 .. 1
__global__ void __launch_bounds__(512) k2_large_tile_merge(const LargeArgs a, uint64_t kk) {
  __shared__ uint32_t tile[LG_TILE];
  const uint64_t n = a.len[blockIdx.y];
  if ((kk >> 1) >= n || a.c[blockIdx.y] == 1) return;
  uint32_t* base = a.tmp + a.off[blockIdx.y];
  for (uint64_t t0 = (uint64_t)blockIdx.x * LG_TILE; t0 < n; t0 += (uint64_t)gridDim.x * LG_TILE) {
    const uint32_t m = (uint32_t)((n - t0) < LG_TILE ? (n - t0) : LG_TILE);
    for (uint32_t i = threadIdx.x; i < m; i += 512) tile[i] = base[t0 + i];
    __syncthreads();
    for (uint32_t j = LG_TILE / 2; j >= 1; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < LG_TILE / 2; t += 512) {
        const uint32_t i = (t / j) * (j << 1) + (t % j);
        const uint32_t l = i + j;
        if (l < m) {
          const uint32_t x = tile[i], y = tile[l];
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

// one CTA per heavy group: in-place dedup + filter, encode, record, bucket totals
__global__ void __launch_bounds__(1024) k2_large_finish(const LargeArgs a) {
  __shared__ uint64_t s_ws[1024 / 32 + 2];
  __shared__ uint32_t s_stage[intcomp::kStageWords];
  __shared__ uint32_t s_enc;
  __shared__ unsigned long long s_eoff;
  const uint32_t g = blockIdx.x;
  const uint64_t n = a.len[g];
  const bool single = a.c[g] == 1;  // pass-through: duplicates stay
  uint32_t* v = a.tmp + a.off[g];
  uint64_t outn = 0;
  for (uint64_t e0 = 0; e0 < n; e0 += 1024) {
    const uint64_t e = e0 + threadIdx.x;
    const bool valid = e < n;
    const uint32_t x = valid ? v[e] : 0u;
    const uint64_t keep =
        (valid && (single || e == 0 || v[e - 1] != x) && !is_removed(a.rem, x)) ? 1u : 0u;
    uint64_t tot;
    const uint64_t ex = block_exclusive_scan(keep, s_ws, tot);  // syncs: reads precede writes
    if (keep) v[outn + ex] = x;
    outn += tot;
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  if (warp_id() == 0) {
    const uint32_t enc = a.want_enc ? intcomp::enc_size_warp(v, (uint32_t)outn) : 0u;
    if (lane_id() == 0) {
      s_enc = enc;
      s_eoff = enc ? atomicAdd(a.enc_alloc, (unsigned long long)enc) : 0ull;
    }
    __syncwarp();
    if (enc) intcomp::enc_emit_warp(v, (uint32_t)outn, a.tmp_enc + s_eoff, s_stage);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    GroupRec& r = a.recs[a.rec[g]];
    r.cnt = (uint32_t)outn;
    r.enc = s_enc;
    r.dec = reinterpret_cast<uint64_t>(v);
    r.eoff = s_eoff;
    if (outn || a.keep_empty) {
      const uint32_t b = a.bucket[g];
      unsigned long long* bo = reinterpret_cast<unsigned long long*>(a.bk_raw);
      atomicAdd(&bo[0ull * a.nb1 + b], 1ull);
      atomicAdd(&bo[1ull * a.nb1 + b], (unsigned long long)r.tlen);
      atomicAdd(&bo[2ull * a.nb1 + b], (unsigned long long)outn);
      atomicAdd(&bo[3ull * a.nb1 + b], (unsigned long long)s_enc);
    }
  }
}

// Σ bk_D (terms_merged)
__global__ void __launch_bounds__(1024)
k12_sum_D(const uint32_t* __restrict__ d, uint32_t n, uint64_t* __restrict__ out) {
  __shared__ uint64_t ws[1024 / 32 + 2];
  uint64_t acc = 0;
  for (uint32_t i = threadIdx.x; i < n; i += 1024) acc += d[i];
  uint64_t tot;
  block_exclusive_scan(acc, ws, tot);
  if (threadIdx.x == 0) *out = tot;
}

// ---------------------------------------------------------------- host driver
int k12_union(const MergePlan& plan, const RemovedSet& rem, bool want_dec, bool want_enc,
              bool keep_empty, uint64_t n_in, UnionOut& u, cudaStream_t s) {
  u.keep_empty = keep_empty;
  u.want_dec = want_dec;
  u.want_enc = want_enc;
  const uint32_t B = plan.n_buckets, N = plan.n_total;
  const int k = plan.k;
  II2_TRY(u.recs.alloc(N, s));
  if (want_dec) II2_TRY(u.tmp_post.alloc(n_in, s));
  const uint64_t enc_cap = n_in + n_in / 4 + 3 * std::min<uint64_t>(N, n_in) + 64;
  if (want_enc) II2_TRY(u.tmp_enc.alloc(enc_cap, s));
  II2_TRY(u.bk_D.alloc(B, s));
  II2_TRY(u.bk_raw.alloc(4 * (size_t)(B + 1), s));
  II2_TRY(u.bk_out.alloc(4 * (size_t)(B + 1), s));
  // totals[0..3] scan totals, [4] terms merged, [5] enc allocator, [6] n_large (u32), [7] n_lsrc
  II2_TRY(u.totals.alloc(8, s));
  const uint32_t large_cap = (uint32_t)(n_in / CAP_P + 1);  // every heavy group holds > CAP_P
  const uint64_t lsrc_cap = std::min<uint64_t>(N, (uint64_t)large_cap * k);
  DevBuf<uint32_t> large_u32;
  DevBuf<uint64_t> lsrc;
  II2_TRY(large_u32.alloc(4 * (size_t)large_cap, s));
  II2_TRY(lsrc.alloc(2 * (size_t)lsrc_cap, s));
  II2_CUDA_TRY(cudaMemsetAsync(u.totals.p, 0, 64, s));
  // row B of bk_raw (the scan's sentinel column) must be zero
  II2_CUDA_TRY(cudaMemsetAsync(u.bk_raw.p, 0, 4 * (size_t)(B + 1) * 8, s));

  K12Args a;
  a.segs = plan.segs;
  a.k = k;
  a.part = plan.part.p;
  a.row_of = plan.row_of.p;
  a.bk_pos = plan.bk_pos();
  a.bk_P = plan.bk_P();
  a.bk_cpl = plan.bk_cpl.p;
  a.rem = rem;
  a.want_enc = want_enc ? 1 : 0;
  a.want_dec = want_dec ? 1 : 0;
  a.keep_empty = keep_empty ? 1 : 0;
  a.recs = u.recs.p;
  a.tmp_post = u.tmp_post.p;
  a.tmp_enc = u.tmp_enc.p;
  a.enc_alloc = reinterpret_cast<unsigned long long*>(u.totals.p + 5);
  a.bk_raw = u.bk_raw.p;
  a.bk_D = u.bk_D.p;
  a.nb1 = B + 1;
  a.n_large = reinterpret_cast<uint32_t*>(u.totals.p + 6);
  a.large_rec = large_u32.p;
  a.large_beg = large_u32.p + large_cap;
  a.large_c = large_u32.p + 2 * (size_t)large_cap;
  a.large_bucket = large_u32.p + 3 * (size_t)large_cap;
  a.n_lsrc = reinterpret_cast<unsigned long long*>(u.totals.p + 7);
  a.lsrc_ptr = lsrc.p;
  a.lsrc_len = lsrc.p + lsrc_cap;

  const size_t smem = k12_smem_bytes(k);
  static size_t attr = 0;
  if (smem > attr) {
    II2_CUDA_TRY(cudaFuncSetAttribute(k12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)k12_smem_bytes(kMaxSegs)));
    attr = k12_smem_bytes(kMaxSegs);
  }
  {
    ProfScope scope("k12_union", s);
    k12_kernel<<<B, K12_THREADS, smem, s>>>(a);
    II2_LAUNCHED();
  }
  k12_sum_D<<<1, 1024, 0, s>>>(u.bk_D.p, B, u.totals.p + 4);
  II2_LAUNCHED();
  // optimistic: scan right away; redone only if heavy groups were deferred
  II2_TRY(exclusive_scan_multi_u64(u.bk_raw.p, u.bk_out.p, B + 1, 4, u.totals.p, s));
  uint64_t h_tot[8];
  II2_CUDA_TRY(cudaMemcpyAsync(h_tot, u.totals.p, 64, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  const uint32_t h_nl = (uint32_t)h_tot[6];
  if (h_nl > 0) {
    ProfScope scope("k2_large", s);
    DevBuf<uint64_t> d_len, d_off;
    II2_TRY(d_len.alloc(h_nl, s));
    II2_TRY(d_off.alloc(h_nl, s));
    LargeArgs la;
    la.rec = a.large_rec;
    la.beg = a.large_beg;
    la.c = a.large_c;
    la.bucket = a.large_bucket;
    la.lsrc_ptr = a.lsrc_ptr;
    la.lsrc_len = a.lsrc_len;
    la.len = d_len.p;
    la.off = d_off.p;
    la.tmp = nullptr;
    la.recs = u.recs.p;
    la.rem = rem;
    la.want_enc = a.want_enc;
    la.keep_empty = a.keep_empty;
    la.tmp_enc = u.tmp_enc.p;
    la.enc_alloc = a.enc_alloc;
    la.bk_raw = u.bk_raw.p;
    la.nb1 = B + 1;
    k2_large_len<<<div_up((uint64_t)h_nl * 32, 256), 256, 0, s>>>(la, h_nl);
    II2_LAUNCHED();
    std::vector<uint64_t> lens(h_nl), offs(h_nl);
    II2_CUDA_TRY(cudaMemcpyAsync(lens.data(), d_len.p, (size_t)h_nl * 8, cudaMemcpyDeviceToHost, s));
    II2_CUDA_TRY(cudaStreamSynchronize(s));
    uint64_t total = 0;
    for (uint32_t i = 0; i < h_nl; i++) {
      if (lens[i] >= (1ull << 32)) {
        set_last_error("a single term unions %llu postings (max 2^32-1)",
                       (unsigned long long)lens[i]);
        return II2_ERR_UNSUPPORTED;
      }
      offs[i] = total;
      total += lens[i];
    }
    II2_TRY(u.large_tmp.alloc(total, s));
    la.tmp = u.large_tmp.p;
    II2_CUDA_TRY(cudaMemcpyAsync(d_off.p, offs.data(), (size_t)h_nl * 8, cudaMemcpyHostToDevice, s));
    for (uint32_t y0 = 0; y0 < h_nl; y0 += 32768) {  // grid.y limit
      const uint32_t ny = std::min<uint32_t>(32768, h_nl - y0);
      uint64_t maxL = 0;
      for (uint32_t i = 0; i < ny; i++) maxL = std::max(maxL, lens[y0 + i]);
      LargeArgs b2 = la;
      b2.rec += y0;
      b2.beg += y0;
      b2.c += y0;
      b2.bucket += y0;
      b2.len += y0;
      b2.off += y0;
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
    II2_CUDA_TRY(cudaMemcpyAsync(h_tot, u.totals.p, 64, cudaMemcpyDeviceToHost, s));
    II2_CUDA_TRY(cudaStreamSynchronize(s));  // also keeps `offs` alive until its copy is done
  }
  for (int i = 0; i < 4; i++) u.h_totals[i] = h_tot[i];
  u.terms_merged = h_tot[4];
  return II2_OK;
}

}  // namespace ii2
