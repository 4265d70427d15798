// k1_term_merge.cu — K1: k-way merge of sorted term dictionaries across segments.
//
// Replaces go-iterators' MergingIterator over per-segment readers (shard.go:253-278,
// ordering = file.CompareTermValues = bytes.Compare, file/types.go:24-26).  Output is the
// MERGED ORDER of all term instances with equal terms grouped (see plan.cuh); the posting
// union (K2) consumes it.  One pass over the term bytes:
//
//   k1_sample_sort   <= 8192 evenly spaced sample terms, sorted in one CTA's shared memory
//                    -> splitters.
//   k1_partition     lower_bound of every splitter in every segment -> bucket table.  A
//                    bucket holds ALL instances of the terms in [splitter_b-1, splitter_b).
//   k1_bucket_pos    bucket start positions + common-prefix length per bucket.
//   k1_merge_tiles   one CTA per bucket.  Oversized buckets are bisected on the fly (pivot =
//                    median of the widest run) into tiles of <= 2048 instances.  Per tile:
//                    16-byte big-endian key windows (past the bucket's common prefix) are
//                    loaded with aligned 32-bit loads; equal terms are grouped with a
//                    shared-memory hash table; only the DISTINCT terms are sorted (bitonic
//                    network); instances are scattered to their merged positions.
//
// Integer/byte work, HBM-bound by design: each term byte and offset is read once from HBM
// (binary-search probes hit L2), each instance writes 18 bytes of plan.
#include "plan.cuh"

namespace ii2 {

constexpr int K1_CAP = 2048;       // instances per tile
constexpr int K1_THREADS = 512;
constexpr int K1_HT = 2 * K1_CAP;  // hash slots
constexpr uint16_t K1_EMPTY = 0xFFFFu;

// ---------------------------------------------------------------- helpers
// First 8 bytes of a term as a big-endian integer, zero padded (bytewise loads; cold path).
__device__ __forceinline__ uint64_t key8_bytes(const uint8_t* t, uint32_t n) {
  uint64_t k = 0;
  for (uint32_t i = 0; i < 8; i++) k = (k << 8) | (i < n ? t[i] : 0);
  return k;
}

// Bytes [c, c+16) of the term at tb+g0 (length len >= c) as two big-endian u64, zero
// padded past the end.  Five aligned 32-bit loads + funnel shifts; tb must be 4-byte
// aligned and readable 20 bytes past the last term byte.
__device__ __forceinline__ void load_key16(const uint8_t* __restrict__ tb, uint32_t g0,
                                           uint32_t len, uint32_t c, uint64_t& hi, uint64_t& lo) {
  const uint32_t avail = len - c;
  if (avail == 0) {
    hi = lo = 0;
    return;
  }
  const uint32_t a = g0 + c;
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(tb + (a & ~3u));
  const uint32_t sh = (a & 3u) * 8u;
  uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3),
           w4 = __ldg(wp + 4);
  uint32_t x0 = __byte_perm(__funnelshift_r(w0, w1, sh), 0, 0x0123);
  uint32_t x1 = __byte_perm(__funnelshift_r(w1, w2, sh), 0, 0x0123);
  uint32_t x2 = __byte_perm(__funnelshift_r(w2, w3, sh), 0, 0x0123);
  uint32_t x3 = __byte_perm(__funnelshift_r(w3, w4, sh), 0, 0x0123);
  hi = ((uint64_t)x0 << 32) | x1;
  lo = ((uint64_t)x2 << 32) | x3;
  if (avail < 16) {
    if (avail <= 8) {
      lo = 0;
      if (avail < 8) hi &= ~0ull << (8 * (8 - avail));
    } else {
      lo &= ~0ull << (8 * (16 - avail));
    }
  }
}

__device__ __forceinline__ uint32_t hash_key(uint64_t hi, uint64_t lo, uint32_t len) {
  uint64_t h = hi * 0x9E3779B97F4A7C15ull;
  h ^= (lo + 0xD6E8FEB86659FD93ull + (h << 6) + (h >> 2));
  h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33;
  h += len * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  return (uint32_t)h;
}

template <typename T>
__device__ __forceinline__ T block_max(T v, T* ws) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    T o = __shfl_xor_sync(0xffffffffu, v, d);
    v = o > v ? o : v;
  }
  __syncthreads();
  if (lane_id() == 0) ws[warp_id()] = v;
  __syncthreads();
  const unsigned nw = (blockDim.x + 31) >> 5;
  T r = ws[0];
  for (unsigned i = 1; i < nw; i++) r = ws[i] > r ? ws[i] : r;
  return r;
}

// ---------------------------------------------------------------- k1_sample_sort
// One CTA.  Sample j is global instance (j+1)*N_T/(S+1).  Sorted ascending by term; the
// result (instance ids) goes to split[0..S).
__global__ void __launch_bounds__(1024)
k1_sample_sort(const SegDesc* __restrict__ segs, int k, uint32_t n_total, uint32_t S,
               uint32_t* __restrict__ split) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* key = reinterpret_cast<uint64_t*>(smem_raw);               // [S]
  uint32_t* inst = reinterpret_cast<uint32_t*>(key + S);               // [S]
  uint16_t* perm = reinterpret_cast<uint16_t*>(inst + S);              // [S]
  for (uint32_t j = threadIdx.x; j < S; j += blockDim.x) {
    uint32_t g = (uint32_t)(((uint64_t)(j + 1) * n_total) / (S + 1));
    int s;
    uint32_t idx;
    locate_instance(segs, k, g, s, idx);
    uint32_t o = segs[s].toff[idx], n = segs[s].toff[idx + 1] - o;
    key[j] = key8_bytes(segs[s].tb + o, n);
    inst[j] = g;
    perm[j] = (uint16_t)j;
  }
  __syncthreads();
  auto less = [&](uint16_t a, uint16_t b) -> bool {
    uint64_t ka = key[a], kb = key[b];
    if (ka != kb) return ka < kb;
    if (inst[a] == inst[b]) return false;
    int sa, sb;
    uint32_t ia, ib;
    locate_instance(segs, k, inst[a], sa, ia);
    locate_instance(segs, k, inst[b], sb, ib);
    uint32_t oa = segs[sa].toff[ia], na = segs[sa].toff[ia + 1] - oa;
    uint32_t ob = segs[sb].toff[ib], nb = segs[sb].toff[ib + 1] - ob;
    int c = term_compare(segs[sa].tb + oa, na, segs[sb].tb + ob, nb);
    if (c) return c < 0;
    return inst[a] < inst[b];
  };
  bitonic_sort_any(perm, S, threadIdx.x, blockDim.x, less, [] { __syncthreads(); });
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < S; j += blockDim.x) split[j] = inst[perm[j]];
}

// ---------------------------------------------------------------- k1_partition
// part[row*k + s]: row 0 = window start, rows 1..S = lower_bound(splitter row-1), row S+1 =
// window end.  One thread per (row, segment).
__global__ void __launch_bounds__(256)
k1_partition(const SegDesc* __restrict__ segs, int k, uint32_t S, const uint32_t* __restrict__ split,
             uint32_t* __restrict__ part) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t total = (uint64_t)(S + 2) * k;
  if (t >= total) return;
  uint32_t row = (uint32_t)(t / k);
  int s = (int)(t % k);
  const SegDesc sd = segs[s];
  uint32_t v;
  if (row == 0) {
    v = sd.lo;
  } else if (row == S + 1) {
    v = sd.hi;
  } else {
    int ss;
    uint32_t idx;
    locate_instance(segs, k, split[row - 1], ss, idx);
    if (ss == s) {
      v = idx;
    } else {
      uint32_t o = segs[ss].toff[idx], n = segs[ss].toff[idx + 1] - o;
      v = seg_lower_bound(sd, sd.lo, sd.hi, segs[ss].tb + o, n);
    }
  }
  part[t] = v;
}

// One warp per row: bk_pos[row] = Σ_s (part[row][s] - lo_s); bucket `row` (< B) also gets the
// common-prefix length of its two delimiting splitters (0 for the open-ended buckets).
__global__ void __launch_bounds__(256)
k1_bucket_pos(const SegDesc* __restrict__ segs, int k, uint32_t S, const uint32_t* __restrict__ split,
              const uint32_t* __restrict__ part, uint32_t* __restrict__ bk_pos,
              uint32_t* __restrict__ bk_cpl) {
  uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row > S + 1) return;
  uint32_t acc = 0;
  for (int s = lane_id(); s < k; s += 32) acc += part[(uint64_t)row * k + s] - segs[s].lo;
  acc = warp_sum(acc);
  if (lane_id() == 0) {
    bk_pos[row] = acc;
    if (row <= S) {  // bucket `row` lies between splitter row-1 and splitter row
      uint32_t c = 0;
      if (row >= 1 && row < S) {
        int sa, sb;
        uint32_t ia, ib;
        locate_instance(segs, k, split[row - 1], sa, ia);
        locate_instance(segs, k, split[row], sb, ib);
        uint32_t oa = segs[sa].toff[ia], na = segs[sa].toff[ia + 1] - oa;
        uint32_t ob = segs[sb].toff[ib], nb = segs[sb].toff[ib + 1] - ob;
        const uint8_t* a = segs[sa].tb + oa;
        const uint8_t* b = segs[sb].tb + ob;
        uint32_t m = na < nb ? na : nb;
        while (c < m && a[c] == b[c]) c++;
      }
      bk_cpl[row] = c;
    }
  }
}

// ---------------------------------------------------------------- k1_merge_tiles
struct TileSmem {
  uint64_t* key_hi;   // [CAP]
  uint64_t* key_lo;   // [CAP]
  uint32_t* tg0;      // [CAP] byte offset of the term inside its segment's term_bytes
  uint32_t* cnt;      // [CAP] group size, indexed by representative
  uint32_t* cursor;   // [CAP]
  uint32_t* cur;      // [K]   tile start per segment
  uint32_t* mm;       // [K]   tile end per segment
  uint32_t* hi;       // [K]   bisection upper bounds
  uint32_t* endr;     // [K]   bucket end per segment
  uint32_t* rstart;   // [K+1] run starts inside the tile
  uint16_t* tlen;     // [CAP]
  uint16_t* tseg;     // [CAP]
  uint16_t* grp;      // [CAP] representative of the instance's group
  uint16_t* gstart;   // [CAP] first tile position of the group, indexed by representative
  uint16_t* reps;     // [CAP] distinct representatives, then sorted by term
  uint16_t* table;    // [HT]
};

__host__ __device__ inline size_t k1_tile_smem_bytes(int K) {
  return (size_t)K1_CAP * (8 + 8 + 4 + 4 + 4) + (size_t)(5 * K + 1) * 4 +
         (size_t)K1_CAP * 2 * 5 + (size_t)K1_HT * 2 + 64;
}

__global__ void __launch_bounds__(K1_THREADS, 2)
k1_merge_tiles(const SegDesc* __restrict__ segs, int k, const uint32_t* __restrict__ part,
               const uint32_t* __restrict__ bk_pos, const uint32_t* __restrict__ bk_cpl,
               uint32_t* __restrict__ ord_inst, uint64_t* __restrict__ src_ptr,
               uint32_t* __restrict__ src_len, uint16_t* __restrict__ gsz,
               uint64_t* __restrict__ bk_P, uint64_t* __restrict__ bk_D) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint64_t s_ws64[K1_THREADS / 32 + 2];
  __shared__ uint32_t s_ws32[K1_THREADS / 32 + 2];
  TileSmem sm;
  {
    uint8_t* p = smem_raw;
    sm.key_hi = reinterpret_cast<uint64_t*>(p); p += K1_CAP * 8;
    sm.key_lo = reinterpret_cast<uint64_t*>(p); p += K1_CAP * 8;
    sm.tg0 = reinterpret_cast<uint32_t*>(p); p += K1_CAP * 4;
    sm.cnt = reinterpret_cast<uint32_t*>(p); p += K1_CAP * 4;
    sm.cursor = reinterpret_cast<uint32_t*>(p); p += K1_CAP * 4;
    sm.cur = reinterpret_cast<uint32_t*>(p); p += k * 4;
    sm.mm = reinterpret_cast<uint32_t*>(p); p += k * 4;
    sm.hi = reinterpret_cast<uint32_t*>(p); p += k * 4;
    sm.endr = reinterpret_cast<uint32_t*>(p); p += k * 4;
    sm.rstart = reinterpret_cast<uint32_t*>(p); p += (k + 1) * 4;
    sm.tlen = reinterpret_cast<uint16_t*>(p); p += K1_CAP * 2;
    sm.tseg = reinterpret_cast<uint16_t*>(p); p += K1_CAP * 2;
    sm.grp = reinterpret_cast<uint16_t*>(p); p += K1_CAP * 2;
    sm.gstart = reinterpret_cast<uint16_t*>(p); p += K1_CAP * 2;
    sm.reps = reinterpret_cast<uint16_t*>(p); p += K1_CAP * 2;
    sm.table = reinterpret_cast<uint16_t*>(p);
  }
  const uint32_t tid = threadIdx.x;
  const uint32_t b = blockIdx.x;
  uint32_t W = bk_pos[b + 1] - bk_pos[b];
  if (W == 0) {
    if (tid == 0) {
      bk_P[b] = 0;
      bk_D[b] = 0;
    }
    return;
  }
  for (int s = tid; s < k; s += K1_THREADS) {
    sm.cur[s] = part[(uint64_t)b * k + s];
    sm.endr[s] = part[(uint64_t)(b + 1) * k + s];
  }
  __syncthreads();
  uint32_t out_base = bk_pos[b];
  const uint32_t cpl = bk_cpl[b];
  uint64_t p_acc = 0;  // per-thread partial of Σ list lengths
  uint32_t d_acc = 0;  // uniform: distinct terms so far

  while (W > 0) {
    // ---------------- choose the tile [cur, mm) ----------------
    uint32_t size;
    if (W <= (uint32_t)K1_CAP) {
      for (int s = tid; s < k; s += K1_THREADS) sm.mm[s] = sm.endr[s];
      size = W;
      __syncthreads();
    } else {
      for (int s = tid; s < k; s += K1_THREADS) sm.hi[s] = sm.endr[s];
      __syncthreads();
      for (;;) {
        // widest run and its median term = pivot
        uint64_t best = 0;
        for (int s = tid; s < k; s += K1_THREADS) {
          uint64_t cand = ((uint64_t)(sm.hi[s] - sm.cur[s]) << 32) | (uint32_t)s;
          best = cand > best ? cand : best;
        }
        best = block_max(best, s_ws64);
        const int sp = (int)(uint32_t)best;
        const uint32_t win = (uint32_t)(best >> 32);  // >= 2 because Σ windows > CAP >= 2k
        const uint32_t mid = sm.cur[sp] + (win >> 1);
        const SegDesc ps = segs[sp];
        const uint32_t po = ps.toff[mid], pn = ps.toff[mid + 1] - po;
        uint32_t part_sum = 0;
        for (int s = tid; s < k; s += K1_THREADS) {
          uint32_t m;
          if (s == sp) {
            m = mid;
          } else {
            const SegDesc sd = segs[s];
            m = seg_lower_bound(sd, sm.cur[s], sm.hi[s], ps.tb + po, pn);
          }
          sm.mm[s] = m;
          part_sum += m - sm.cur[s];
        }
        uint32_t tot;
        block_exclusive_scan(part_sum, s_ws32, tot);
        size = tot;  // >= 1: the pivot's own run contributes mid - cur >= 1
        if (size <= (uint32_t)K1_CAP) break;
        for (int s = tid; s < k; s += K1_THREADS) sm.hi[s] = sm.mm[s];
        __syncthreads();
      }
      __syncthreads();
    }

    // ---------------- (1) run starts ----------------
    {
      // k <= 1024 = 2 * K1_THREADS: each thread owns segments 2t and 2t+1
      int s0 = 2 * tid, s1 = s0 + 1;
      uint32_t v0 = s0 < k ? sm.mm[s0] - sm.cur[s0] : 0;
      uint32_t v1 = s1 < k ? sm.mm[s1] - sm.cur[s1] : 0;
      uint32_t tot;
      uint32_t ex = block_exclusive_scan(v0 + v1, s_ws32, tot);
      if (s0 < k) sm.rstart[s0] = ex;
      if (s1 < k) sm.rstart[s1] = ex + v0;
      if (tid == 0) sm.rstart[k] = size;
    }
    for (uint32_t t = tid; t < (uint32_t)K1_HT; t += K1_THREADS) sm.table[t] = K1_EMPTY;
    __syncthreads();

    // ---------------- (2) key windows ----------------
    for (uint32_t i = tid; i < size; i += K1_THREADS) {
      int lo = 0, hi = k;  // first s with rstart[s+1] > i
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sm.rstart[mid + 1] <= i)
          lo = mid + 1;
        else
          hi = mid;
      }
      const int s = lo;
      const uint32_t idx = sm.cur[s] + (i - sm.rstart[s]);
      const uint8_t* tb = segs[s].tb;
      const uint32_t* toff = segs[s].toff;
      const uint32_t o = __ldg(toff + idx), n = __ldg(toff + idx + 1) - o;
      uint64_t kh, kl;
      load_key16(tb, o, n, cpl, kh, kl);
      sm.key_hi[i] = kh;
      sm.key_lo[i] = kl;
      sm.tlen[i] = (uint16_t)n;
      sm.tg0[i] = o;
      sm.tseg[i] = (uint16_t)s;
      sm.cnt[i] = 0;
      sm.cursor[i] = 0;
    }
    __syncthreads();

    // bytes past the 16-byte window, only needed for terms longer than cpl+16
    auto tail_compare = [&](uint32_t a, uint32_t bb) -> int {
      const uint32_t skip = cpl + 16;
      const uint32_t na = sm.tlen[a], nb = sm.tlen[bb];
      if (na > skip && nb > skip) {
        const uint8_t* pa = segs[sm.tseg[a]].tb + sm.tg0[a] + skip;
        const uint8_t* pb = segs[sm.tseg[bb]].tb + sm.tg0[bb] + skip;
        return term_compare(pa, na - skip, pb, nb - skip);
      }
      return na < nb ? -1 : (na > nb ? 1 : 0);
    };

    // ---------------- (3) group equal terms (hash table of representatives) ----------------
    for (uint32_t i = tid; i < size; i += K1_THREADS) {
      const uint64_t kh = sm.key_hi[i], kl = sm.key_lo[i];
      uint32_t slot = hash_key(kh, kl, sm.tlen[i]) & (K1_HT - 1);
      uint32_t rep;
      for (;;) {
        unsigned short prev = atomicCAS(reinterpret_cast<unsigned short*>(&sm.table[slot]),
                                        (unsigned short)K1_EMPTY, (unsigned short)i);
        if (prev == K1_EMPTY) {
          rep = i;
          break;
        }
        if (sm.key_hi[prev] == kh && sm.key_lo[prev] == kl && sm.tlen[prev] == sm.tlen[i] &&
            tail_compare(i, prev) == 0) {
          rep = prev;
          break;
        }
        slot = (slot + 1) & (K1_HT - 1);
      }
      sm.grp[i] = (uint16_t)rep;
      atomicAdd(&sm.cnt[rep], 1u);
    }
    __syncthreads();

    // ---------------- (4) list of distinct representatives ----------------
    uint32_t D = 0;
    for (uint32_t base = 0; base < size; base += K1_THREADS) {
      uint32_t i = base + tid;
      uint32_t f = (i < size && sm.grp[i] == i) ? 1u : 0u;
      uint32_t tot;
      uint32_t ex = block_exclusive_scan(f, s_ws32, tot);
      if (f) sm.reps[D + ex] = (uint16_t)i;
      D += tot;
    }
    __syncthreads();

    // ---------------- (5) sort the distinct terms ----------------
    auto less = [&](uint16_t a, uint16_t bb) -> bool {
      uint64_t ha = sm.key_hi[a], hb = sm.key_hi[bb];
      if (ha != hb) return ha < hb;
      uint64_t la = sm.key_lo[a], lb = sm.key_lo[bb];
      if (la != lb) return la < lb;
      return tail_compare(a, bb) < 0;
    };
    bitonic_sort_any(sm.reps, D, tid, (uint32_t)K1_THREADS, less, [] { __syncthreads(); });
    __syncthreads();

    // ---------------- (6) group start positions, in sorted order ----------------
    {
      uint32_t run = 0;
      for (uint32_t base = 0; base < D; base += K1_THREADS) {
        uint32_t r = base + tid;
        uint32_t c = r < D ? sm.cnt[sm.reps[r]] : 0;
        uint32_t tot;
        uint32_t ex = block_exclusive_scan(c, s_ws32, tot);
        if (r < D) sm.gstart[sm.reps[r]] = (uint16_t)(run + ex);
        run += tot;
      }
    }
    __syncthreads();

    // ---------------- (7) scatter instances to their merged positions ----------------
    for (uint32_t i = tid; i < size; i += K1_THREADS) {
      const uint32_t g = sm.grp[i];
      const uint32_t slot = atomicAdd(&sm.cursor[g], 1u);
      const uint32_t P = out_base + sm.gstart[g] + slot;
      const int s = sm.tseg[i];
      const uint32_t idx = sm.cur[s] + (i - sm.rstart[s]);
      const SegDesc sd = segs[s];
      const uint64_t p0 = __ldg(sd.poff + idx), p1 = __ldg(sd.poff + idx + 1);
      ord_inst[P] = sd.base + (idx - sd.lo);
      src_ptr[P] = reinterpret_cast<uint64_t>(sd.post + p0);
      src_len[P] = (uint32_t)(p1 - p0);
      gsz[P] = slot == 0 ? (uint16_t)sm.cnt[g] : (uint16_t)0;
      p_acc += p1 - p0;
    }
    d_acc += D;
    out_base += size;
    W -= size;
    __syncthreads();
    for (int s = tid; s < k; s += K1_THREADS) sm.cur[s] = sm.mm[s];
    __syncthreads();
  }

  uint64_t ptot;
  block_exclusive_scan(p_acc, s_ws64, ptot);
  if (tid == 0) {
    bk_P[b] = ptot;
    bk_D[b] = d_acc;
  }
}

// ---------------------------------------------------------------- host driver
int k1_build_plan(MergePlan& plan, cudaStream_t s) {
  const int k = plan.k;
  const uint32_t N = plan.n_total;
  if (k > kMaxSegs) {
    set_last_error("k1: %d segments in one pass (max %d)", k, kMaxSegs);
    return II2_ERR_UNSUPPORTED;
  }
  // samples: aim at ~one tile per bucket, at most kMaxSamples (sorted by one CTA)
  uint32_t S = N / (K1_CAP / 2);
  if (S > (uint32_t)kMaxSamples) S = kMaxSamples;
  const uint32_t B = S + 1;
  plan.n_buckets = B;
  II2_TRY(plan.part.alloc((size_t)(B + 1) * k, s));
  II2_TRY(plan.bk_pos.alloc(B + 1, s));
  II2_TRY(plan.bk_cpl.alloc(B, s));
  II2_TRY(plan.ord_inst.alloc(N, s));
  II2_TRY(plan.src_ptr.alloc(N, s));
  II2_TRY(plan.src_len.alloc(N, s));
  II2_TRY(plan.gsz.alloc(N, s));
  II2_TRY(plan.bk_PD.alloc(2 * (size_t)(B + 1), s));
  II2_TRY(plan.totals.alloc(2, s));
  DevBuf<uint32_t> split;
  II2_TRY(split.alloc(S ? S : 1, s));

  ProfScope split_scope("k1_split", s);
  if (S > 0) {
    size_t smem = (size_t)S * (8 + 4 + 2) + 16;
    static bool attr1 = false;
    if (!attr1) {
      II2_CUDA_TRY(cudaFuncSetAttribute(k1_sample_sort, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kMaxSamples * 14 + 16));
      attr1 = true;
    }
    k1_sample_sort<<<1, 1024, smem, s>>>(plan.segs, k, N, S, split.p);
    II2_LAUNCHED();
  }
  {
    uint64_t total = (uint64_t)(S + 2) * k;
    k1_partition<<<div_up(total, 256), 256, 0, s>>>(plan.segs, k, S, split.p, plan.part.p);
    II2_LAUNCHED();
    k1_bucket_pos<<<div_up((uint64_t)(S + 2) * 32, 256), 256, 0, s>>>(
        plan.segs, k, S, split.p, plan.part.p, plan.bk_pos.p, plan.bk_cpl.p);
    II2_LAUNCHED();
  }
  {
    size_t smem = k1_tile_smem_bytes(k);
    static size_t attr2 = 0;
    if (smem > attr2) {
      II2_CUDA_TRY(cudaFuncSetAttribute(k1_merge_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)k1_tile_smem_bytes(kMaxSegs)));
      attr2 = k1_tile_smem_bytes(kMaxSegs);
    }
    II2_CUDA_TRY(cudaMemsetAsync(plan.bk_P() + B, 0, 8, s));
    II2_CUDA_TRY(cudaMemsetAsync(plan.bk_D() + B, 0, 8, s));
    split_scope.end();
    ProfScope scope("k1_merge_tiles", s);
    k1_merge_tiles<<<B, K1_THREADS, smem, s>>>(plan.segs, k, plan.part.p, plan.bk_pos.p,
                                               plan.bk_cpl.p, plan.ord_inst.p, plan.src_ptr.p,
                                               plan.src_len.p, plan.gsz.p, plan.bk_P(), plan.bk_D());
    II2_LAUNCHED();
  }
  II2_TRY(exclusive_scan_multi_u64(plan.bk_PD.p, plan.bk_PD.p, B + 1, 2, plan.totals.p, s));
  return II2_OK;
}

}  // namespace ii2
