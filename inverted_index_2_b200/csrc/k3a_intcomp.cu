// k3a_intcomp.cu — batched posting codec kernels (K3a): many independent lists per launch.
// Replaces one intcomp.CompressUint32 call per term (file/writer.go:49) and one
// intcomp.UncompressUint32 per term (file/reader.go:100).
//   size pass -> exclusive scan (so `_val` offsets equal the reference's running
//   valuesOffset, file/writer.go:56) -> emit pass.
// Lists below one 128-block are handled one per thread; longer lists one per warp
// (32-lane groups match the codec's 32-value groups); lists of >= 8192 values block-parallel
// over many CTAs (decode: speculative tile parse + per-block sums; encode: slices of blocks).
#include <algorithm>

#include "codec.cuh"
#include "intcomp.cuh"
#include "runtime.cuh"

namespace ii2 {

constexpr int kCodecThreads = 256;

// ---------------------------------------------------------------- decode
// A list whose first bin-pack section holds at least kHugeValues values is decoded block-parallel
// (below); n_work = {warp-per-list entries, huge lists, their blocks, their tiles}.
constexpr uint32_t kHugeValues = 128u * 64u;
constexpr uint32_t kWalkThreads = 1024;
constexpr uint32_t kTileWords = 16384;  // stream words per tile of the header walk
constexpr uint32_t kTileEntries = 512;  // a block is <= 509 words: a tile is entered at offset 0 .. 508
constexpr uint32_t kChainBatch = 16;    // tiles whose maps the chain kernel stages at once

__global__ void __launch_bounds__(kCodecThreads)
k_dec_count(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff, uint64_t nlists,
            uint64_t* __restrict__ counts, uint32_t* __restrict__ n_work,
            uint32_t* __restrict__ huge_list, uint32_t* __restrict__ huge_gstart,
            uint32_t* __restrict__ huge_tstart, int* __restrict__ err) {
  pdl_enter();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlists) return;
  uint64_t a = woff[i], b = woff[i + 1];
  long long c = b >= a ? intcomp::dec_count(words + a, b - a) : -1;
  if (c < 0) {
    atomicExch(err, 1);
    c = 0;
  } else if (b > a && words[a] >= kHugeValues) {
    const uint32_t h = atomicAdd(&n_work[1], 1u);
    huge_list[h] = (uint32_t)i;
    huge_gstart[h] = atomicAdd(&n_work[2], words[a] >> 7);
    huge_tstart[h] = atomicAdd(&n_work[3], (words[a + 1] - 3 + kTileWords - 1) / kTileWords);
  }
  counts[i] = (uint64_t)c;
}

// ---- long lists, block-parallel --------------------------------------------------------
// dec_warp walks a list block by block: the place of a block is known only when the widths of
// the block before it are, and its first value only when every delta before it is summed —
// two dependent chains of global loads, ~3 us per 128 values, which is all there is to
// overlap when a batch holds one 16 M-value list (C4).  Here both chains are cut.
// Where the headers are (speculative parse over fixed tiles of kTileWords stream words):
//   tile maps  a block is at most 509 words, so the header chain enters a tile at one of 509
//              offsets.  One CTA per tile stages "where the block would end if this word were a
//              header" for every word, and 509 threads walk the tile from every possible entry
//              at once: map[entry] = (headers met, offset at which the next tile is entered);
//   chain      one CTA per list composes the maps tile after tile (a shared-memory lookup per
//              tile): the true entry and the first block number of every tile;
//   tile walk  one CTA per tile walks it once more from its true entry and writes the header
//              positions out.
//   (one thread hopping through the whole stream: ~60 clocks per block, 4.2 ms for 16 M values;
//   pointer doubling inside a window: W log W shared-memory traffic, 1.4 .. 4.9 ms.)
// What the first values are:
//   sums       one warp per block: the sum of its deltas;  exclusive scan over all blocks;
//   blocks     one warp per block decodes it from (first value + the deltas before it).
// What follows the first section (the var-byte tail) goes through dec_warp.
__global__ void __launch_bounds__(256)
k_dec_tile_fill(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
                const uint32_t* __restrict__ huge_list, const uint32_t* __restrict__ huge_tstart,
                uint32_t* __restrict__ tile_list) {
  pdl_enter();
  const uint32_t L = blockIdx.x;
  const uint64_t a = woff[huge_list[L]];
  const uint32_t nt = (words[a + 1] - 3 + kTileWords - 1) / kTileWords;
  for (uint32_t t = threadIdx.x; t < nt; t += blockDim.x) tile_list[huge_tstart[L] + t] = L;
}

// stages nxt[] of one tile (dynamic shared memory, kTileWords x 16 bit); returns its length
__device__ __forceinline__ uint32_t dec_tile_stage(const uint32_t* __restrict__ w, uint64_t len,
                                                   uint32_t t, uint16_t* nxt) {
  const uint64_t t0 = 3 + (uint64_t)t * kTileWords;
  const uint32_t wn = (uint32_t)min((uint64_t)kTileWords, len - t0);
  for (uint32_t i = threadIdx.x; i < wn; i += blockDim.x) {
    const uint32_t x = w[t0 + i];
    nxt[i] = (uint16_t)(i + 1 + ((x >> 24) & 0x7Fu) + ((x >> 16) & 0x7Fu) + ((x >> 8) & 0x7Fu) +
                        (x & 0x7Fu));
  }
  __syncthreads();
  return wn;
}

__global__ void __launch_bounds__(kWalkThreads)
k_dec_tile_maps(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
                const uint32_t* __restrict__ huge_list, const uint32_t* __restrict__ huge_tstart,
                const uint32_t* __restrict__ tile_list, uint32_t* __restrict__ maps) {
  pdl_enter();
  extern __shared__ __align__(16) uint16_t nxt_dyn[];
  const uint32_t tile = blockIdx.x, L = tile_list[tile];
  const uint32_t* w = words + woff[huge_list[L]];
  const uint32_t wn = dec_tile_stage(w, w[1], tile - huge_tstart[L], nxt_dyn);
  const uint32_t e = threadIdx.x;
  if (e < kTileEntries) {
    uint32_t p = e, c = 0;
    while (p < wn) {
      p = nxt_dyn[p];
      c++;
    }
    maps[(uint64_t)tile * kTileEntries + e] = (c << 16) | (p - wn);  // c <= 16384, exit <= 508
  }
}

__global__ void __launch_bounds__(kWalkThreads)
k_dec_tile_chain(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
                 const uint32_t* __restrict__ huge_list, const uint32_t* __restrict__ huge_gstart,
                 const uint32_t* __restrict__ huge_tstart, const uint32_t* __restrict__ maps,
                 uint32_t* __restrict__ tile_entry, uint32_t* __restrict__ tile_bbase,
                 uint64_t* __restrict__ bpos, uint32_t* __restrict__ blist, int* __restrict__ err) {
  pdl_enter();
  __shared__ uint32_t smap[kChainBatch * kTileEntries];
  __shared__ uint32_t s_entry, s_bb;
  const uint32_t L = blockIdx.x;
  const uint64_t a = woff[huge_list[L]];
  const uint32_t nb = words[a] >> 7;
  const uint32_t nt = (words[a + 1] - 3 + kTileWords - 1) / kTileWords;
  const uint32_t ts = huge_tstart[L];
  if (threadIdx.x == 0) {
    s_entry = 0;
    s_bb = 0;
  }
  __syncthreads();
  for (uint32_t t0 = 0; t0 < nt; t0 += kChainBatch) {
    const uint32_t m = min(kChainBatch, nt - t0);
    for (uint32_t i = threadIdx.x; i < m * kTileEntries; i += blockDim.x)
      smap[i] = maps[(uint64_t)(ts + t0) * kTileEntries + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t entry = s_entry, bb = s_bb;
      for (uint32_t j = 0; j < m; j++) {
        const uint32_t x = smap[j * kTileEntries + entry];
        tile_entry[ts + t0 + j] = entry;
        tile_bbase[ts + t0 + j] = bb;
        bb += x >> 16;
        entry = x & 0xFFFFu;
      }
      s_entry = entry;
      s_bb = bb;
    }
    __syncthreads();
  }
  // fewer headers than blocks: the section ends before its blocks do (more: trailing words the
  // reference decoder skips as well)
  const uint32_t found = s_bb;
  if (found < nb) {
    if (threadIdx.x == 0) atomicExch(err, 1);
    const uint32_t gs = huge_gstart[L];
    for (uint32_t b = found + threadIdx.x; b < nb; b += blockDim.x) {
      bpos[gs + b] = ~0ull;
      blist[gs + b] = L;
    }
  }
}

__global__ void __launch_bounds__(kWalkThreads)
k_dec_tile_walk(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
                const uint32_t* __restrict__ huge_list, const uint32_t* __restrict__ huge_gstart,
                const uint32_t* __restrict__ huge_tstart, const uint32_t* __restrict__ tile_list,
                const uint32_t* __restrict__ tile_entry, const uint32_t* __restrict__ tile_bbase,
                uint64_t* __restrict__ bpos, uint32_t* __restrict__ blist) {
  pdl_enter();
  extern __shared__ __align__(16) uint16_t nxt_dyn[];
  uint16_t* hdr = nxt_dyn + kTileWords;  // the headers of the tile, in order
  __shared__ uint32_t s_cnt;
  const uint32_t tile = blockIdx.x, L = tile_list[tile];
  const uint64_t a = woff[huge_list[L]];
  const uint32_t* w = words + a;
  const uint32_t t = tile - huge_tstart[L];
  const uint32_t wn = dec_tile_stage(w, w[1], t, nxt_dyn);
  const uint32_t nb = w[0] >> 7, bb = tile_bbase[tile];
  if (threadIdx.x == 0) {
    const uint32_t room = bb < nb ? nb - bb : 0u;
    uint32_t p = tile_entry[tile], c = 0;
    while (p < wn && c < room) {
      hdr[c++] = (uint16_t)p;
      p = nxt_dyn[p];
    }
    s_cnt = c;
  }
  __syncthreads();
  const uint32_t cnt = s_cnt, gs = huge_gstart[L];
  const uint64_t t0 = a + 3 + (uint64_t)t * kTileWords;
  for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) {
    bpos[gs + bb + k] = t0 + hdr[k];
    blist[gs + bb + k] = L;
  }
}

__global__ void __launch_bounds__(kCodecThreads)
k_dec_blocksum(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
               const uint32_t* __restrict__ huge_list, const uint64_t* __restrict__ bpos,
               const uint32_t* __restrict__ blist, uint32_t nblocks, uint64_t* __restrict__ bsum,
               int* __restrict__ err) {
  pdl_enter();
  const uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (blk > nblocks) return;
  uint32_t sum = 0;
  if (blk < nblocks && bpos[blk] != ~0ull) {
    const uint64_t a = woff[huge_list[blist[blk]]];
    bool bad = false;
    sum = intcomp::dec_block_warp<false>(words, bpos[blk], a + words[a + 1], 0u, nullptr, &bad);
    if (bad) {
      if (lane_id() == 0) atomicExch(err, 1);
      sum = 0;
    }
  }
  if (lane_id() == 0) bsum[blk] = sum;  // [nblocks] = 0: room for the scan's total
}

__global__ void __launch_bounds__(kCodecThreads)
k_dec_blocks(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
             const uint32_t* __restrict__ huge_list, const uint32_t* __restrict__ huge_gstart,
             const uint64_t* __restrict__ bpos, const uint32_t* __restrict__ blist,
             uint32_t nblocks, const uint64_t* __restrict__ bscan,
             const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out) {
  pdl_enter();
  const uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (blk >= nblocks || bpos[blk] == ~0ull) return;
  const uint32_t L = blist[blk], gs = huge_gstart[L], i = huge_list[L];
  const uint64_t a = woff[i];
  // sums are taken modulo 2^32 (unsorted lists have negative deltas): the low half of the
  // 64-bit scan is the 32-bit sum
  const uint32_t prev = words[a + 2] + (uint32_t)(bscan[blk] - bscan[gs]);
  bool bad = false;
  intcomp::dec_block_warp<true>(words, bpos[blk], a + words[a + 1], prev,
                                out + out_off[i] + 128ull * (blk - gs), &bad);
}

// the sections after the first one of every long list (normally the var-byte tail)
__global__ void __launch_bounds__(kCodecThreads)
k_dec_huge_tail(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
                const uint32_t* __restrict__ huge_list, uint32_t n_huge,
                const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out,
                int* __restrict__ err) {
  pdl_enter();
  const uint32_t L = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (L >= n_huge) return;
  const uint32_t i = huge_list[L];
  const uint64_t a = woff[i], b = woff[i + 1];
  const uint64_t c0 = words[a], len = words[a + 1];
  if (a + len >= b) return;
  const int rc = intcomp::dec_warp(words + a + len, b - a - len, out + out_off[i] + c0,
                                   out_off[i + 1] - out_off[i] - c0);
  if (rc && lane_id() == 0) atomicExch(err, 1);
}

// FST outputs (byte offsets) + file size -> word offsets [n+1] (file/reader.go:52,64)
__global__ void __launch_bounds__(kCodecThreads)
k_valoff_to_woff(const uint64_t* __restrict__ val_off, uint64_t n, uint64_t val_size,
                 uint64_t* __restrict__ woff, int* __restrict__ err) {
  pdl_enter();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint64_t o = i < n ? val_off[i] : val_size;
  if ((o & 3) || o > val_size) atomicExch(err, 1);
  woff[i] = o >> 2;
}

__global__ void __launch_bounds__(kCodecThreads)
k_dec_short(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff, uint64_t nlists,
            const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out,
            uint32_t* __restrict__ worklist, uint32_t* __restrict__ n_work,
            int* __restrict__ err) {
  pdl_enter();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlists) return;
  uint64_t a = woff[i], b = woff[i + 1];
  if (b <= a) return;  // empty run -> empty list (file/writer_test.go:15)
  if (words[a] >= kHugeValues) return;  // block-parallel path
  if (words[a] >= 128) {
    worklist[atomicAdd(n_work, 1u)] = (uint32_t)i;
    return;
  }
  if (intcomp::dec_varbyte_thread(words + a, b - a, out + out_off[i])) atomicExch(err, 1);
}

__global__ void __launch_bounds__(kCodecThreads)
k_dec_long(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff,
           const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out,
           const uint32_t* __restrict__ worklist, const uint32_t* __restrict__ n_work,
           int* __restrict__ err) {
  pdl_enter();
  const uint32_t nw = *n_work;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nw; t += warps) {
    uint32_t i = worklist[t];
    uint64_t a = woff[i], b = woff[i + 1];
    int rc = intcomp::dec_warp(words + a, b - a, out + out_off[i], out_off[i + 1] - out_off[i]);
    if (rc && lane_id() == 0) atomicExch(err, 1);
  }
}

int intcomp_decode_dev(const uint32_t* d_words, const uint64_t* d_woff, uint64_t nlists,
                       DevBuf<uint32_t>& out, DevBuf<uint64_t>& out_off, uint64_t* total_out,
                       cudaStream_t s, bool scratch_out) {
  ProfScope scope("k3a_decode", s);
  if (scratch_out) {
    II2_TRY(out_off.alloc_scratch(nlists + 1, s));
  } else {
    II2_TRY(out_off.alloc(nlists + 1, s));
  }
  if (nlists >= (1ull << 32)) {
    set_last_error("more than 2^32 lists in one decode batch");
    return II2_ERR_UNSUPPORTED;
  }
  DevBuf<int> err;
  DevBuf<uint32_t> n_work, worklist, huge_list, huge_gstart, huge_tstart;
  DevBuf<uint64_t> d_total;
  II2_TRY(err.alloc_scratch(1, s));
  II2_TRY(n_work.alloc_scratch(4, s));
  II2_TRY(d_total.alloc_scratch(1, s));
  II2_TRY(worklist.alloc_scratch(nlists ? nlists : 1, s));
  II2_TRY(huge_list.alloc_scratch(nlists ? nlists : 1, s));
  II2_TRY(huge_gstart.alloc_scratch(nlists ? nlists : 1, s));
  II2_TRY(huge_tstart.alloc_scratch(nlists ? nlists : 1, s));
  II2_CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
  II2_CUDA_TRY(cudaMemsetAsync(n_work.p, 0, 4 * sizeof(uint32_t), s));
  II2_CUDA_TRY(cudaMemsetAsync(out_off.p + nlists, 0, sizeof(uint64_t), s));
  if (nlists) {
    II2_LAUNCH_CHAIN(k_dec_count, div_up(nlists, kCodecThreads), kCodecThreads, 0, s, d_words, d_woff, nlists, out_off.p, n_work.p, huge_list.p, huge_gstart.p, huge_tstart.p, err.p);
  }
  II2_TRY(exclusive_scan_u64(out_off.p, nlists + 1, d_total.p, s));
  uint64_t total = 0;
  int herr = 0;
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 20, d_total.p, sizeof(total), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 21, err.p, sizeof(herr), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 22, n_work.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  total = pinned_scratch()[20];
  herr = *reinterpret_cast<const int*>(pinned_scratch() + 21);
  const uint32_t n_huge = reinterpret_cast<const uint32_t*>(pinned_scratch() + 22)[1];
  const uint32_t n_hblocks = reinterpret_cast<const uint32_t*>(pinned_scratch() + 22)[2];
  const uint32_t n_htiles = reinterpret_cast<const uint32_t*>(pinned_scratch() + 22)[3];
  if (herr) {
    set_last_error("undecodable intcomp stream in batch");
    return II2_ERR_CORRUPT;
  }
  if (scratch_out) {
    II2_TRY(out.alloc_scratch(total, s, 16));
  } else {
    II2_TRY(out.alloc(total, s, 16));
  }
  if (nlists) {
    II2_LAUNCH_CHAIN(k_dec_short, div_up(nlists, kCodecThreads), kCodecThreads, 0, s, d_words, d_woff, nlists, out_off.p, out.p, worklist.p, n_work.p, err.p);
    II2_LAUNCH_CHAIN(k_dec_long, kNumSMs * 4, kCodecThreads, 0, s, d_words, d_woff, out_off.p, out.p, worklist.p, n_work.p, err.p);
  }
  if (n_huge) {
    DevBuf<uint64_t> bpos, bsum;
    DevBuf<uint32_t> blist;
    II2_TRY(bpos.alloc_scratch(n_hblocks, s));
    II2_TRY(blist.alloc_scratch(n_hblocks, s));
    II2_TRY(bsum.alloc_scratch((size_t)n_hblocks + 1, s));
    {
      DevBuf<uint32_t> tile_list, tile_entry, tile_bbase, maps;
      II2_TRY(tile_list.alloc_scratch(n_htiles + 1, s));
      II2_TRY(tile_entry.alloc_scratch(n_htiles + 1, s));
      II2_TRY(tile_bbase.alloc_scratch(n_htiles + 1, s));
      II2_TRY(maps.alloc_scratch(((size_t)n_htiles + 1) * kTileEntries, s));
      static bool attr_set = false;
      if (!attr_set) {
        II2_CUDA_TRY(cudaFuncSetAttribute(k_dec_tile_walk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(kTileWords * 4)));
        attr_set = true;
      }
      II2_LAUNCH_CHAIN(k_dec_tile_fill, n_huge, 256, 0, s, d_words, d_woff, huge_list.p, huge_tstart.p, tile_list.p);
      if (n_htiles) {
        II2_LAUNCH_CHAIN(k_dec_tile_maps, n_htiles, kWalkThreads, kTileWords * 2, s, d_words, d_woff, huge_list.p, huge_tstart.p, tile_list.p, maps.p);
      }
      II2_LAUNCH_CHAIN(k_dec_tile_chain, n_huge, kWalkThreads, 0, s, d_words, d_woff, huge_list.p, huge_gstart.p, huge_tstart.p, maps.p, tile_entry.p, tile_bbase.p, bpos.p, blist.p, err.p);
      if (n_htiles) {
        II2_LAUNCH_CHAIN(k_dec_tile_walk, n_htiles, kWalkThreads, kTileWords * 4, s, d_words, d_woff, huge_list.p, huge_gstart.p, huge_tstart.p, tile_list.p, tile_entry.p, tile_bbase.p, bpos.p, blist.p);
      }
    }
    II2_LAUNCH_CHAIN(k_dec_blocksum, div_up(((uint64_t)n_hblocks + 1) * 32, kCodecThreads), kCodecThreads, 0, s, d_words, d_woff, huge_list.p, bpos.p, blist.p, n_hblocks, bsum.p, err.p);
    II2_TRY(exclusive_scan_u64(bsum.p, (uint64_t)n_hblocks + 1, nullptr, s));
    II2_LAUNCH_CHAIN(k_dec_blocks, div_up((uint64_t)n_hblocks * 32, kCodecThreads), kCodecThreads, 0, s, d_words, d_woff, huge_list.p, huge_gstart.p, bpos.p, blist.p, n_hblocks, bsum.p, out_off.p, out.p);
    II2_LAUNCH_CHAIN(k_dec_huge_tail, div_up((uint64_t)n_huge * 32, kCodecThreads), kCodecThreads, 0, s, d_words, d_woff, huge_list.p, n_huge, out_off.p, out.p, err.p);
  }
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 24, err.p, sizeof(herr), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  herr = *reinterpret_cast<const int*>(pinned_scratch() + 24);
  if (herr) {
    set_last_error("undecodable intcomp stream in batch");
    return II2_ERR_CORRUPT;
  }
  if (total_out) *total_out = total;
  return II2_OK;
}

int val_offsets_to_word_offsets(const uint64_t* d_val_off, uint64_t n, uint64_t val_size,
                                DevBuf<uint64_t>& woff, cudaStream_t s) {
  II2_TRY(woff.alloc_scratch(n + 1, s));
  DevBuf<int> err;
  II2_TRY(err.alloc_scratch(1, s));
  II2_CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
  II2_LAUNCH_CHAIN(k_valoff_to_woff, div_up(n + 1, kCodecThreads), kCodecThreads, 0, s, d_val_off, n, val_size, woff.p, err.p);
  // (readbacks land in the thread's pinned scratch: a pageable destination would make the copy
  // wait for every transfer in flight on other streams)
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 25, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  const int herr = *reinterpret_cast<const int*>(pinned_scratch() + 25);
  if (herr) {
    set_last_error("_val offsets are not 4-byte aligned or exceed the file size");
    return II2_ERR_CORRUPT;
  }
  return II2_OK;
}

// ---------------------------------------------------------------- encode
constexpr uint64_t kHugeList = 8192;  // values from which a list is encoded by CTAs over slices of its blocks

// n_work[0] counts the warp-per-list entries (filled from the front of `worklist`), n_work[1]
// the CTA-per-list entries (filled from the back)
__global__ void __launch_bounds__(kCodecThreads)
k_enc_size_short(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off, uint64_t nlists,
                 uint64_t* __restrict__ sizes, uint32_t* __restrict__ worklist,
                 uint32_t* __restrict__ n_work) {
  pdl_enter();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlists) return;
  uint64_t a = off[i], n = off[i + 1] - a;
  if (n >= kHugeList) {
    worklist[nlists - 1 - atomicAdd(n_work + 1, 1u)] = (uint32_t)i;
    return;
  }
  if (n >= 128) {
    worklist[atomicAdd(n_work, 1u)] = (uint32_t)i;
    return;
  }
  sizes[i] = intcomp::enc_size_thread_small(in + a, (uint32_t)n);
}

// Huge lists (>= kHugeList values) are cut into slices of blocks, one CTA per slice (grid.y):
// a 16 M-value list is 131 072 independent blocks, not one CTA's work (C4: 14 ms -> well under
// one).  Size pass: block sizes into the list's table, one partial sum per slice; a warp per
// list adds the partials and the tail.  Emit pass: a slice starts after the partials before it.
constexpr uint32_t kSliceBlocks = 64;  // blocks a slice holds at least
__device__ __forceinline__ uint32_t huge_slices(uint32_t nb, uint32_t ymax) {
  const uint32_t y = nb / kSliceBlocks;
  return y < 1 ? 1 : (y > ymax ? ymax : y);
}

__global__ void __launch_bounds__(1024)
k_enc_size_huge(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off, uint64_t nlists,
                uint32_t* __restrict__ tables, uint64_t* __restrict__ part, uint32_t ymax,
                const uint32_t* __restrict__ worklist, const uint32_t* __restrict__ n_work) {
  pdl_enter();
  __shared__ uint64_t ws[1024 / 32 + 2];
  if (blockIdx.x >= n_work[1]) return;
  const uint32_t i = worklist[nlists - 1 - blockIdx.x];
  const uint32_t* v = in + off[i];
  const uint32_t n = (uint32_t)(off[i + 1] - off[i]);
  const uint32_t nb = n >> 7;
  const uint32_t Y = huge_slices(nb, ymax), y = blockIdx.y;
  if (y >= Y) return;
  const uint32_t b0 = (uint32_t)((uint64_t)nb * y / Y), b1 = (uint32_t)((uint64_t)nb * (y + 1) / Y);
  uint32_t* table = tables + ((off[i] - off[0]) >> 7);
  uint64_t acc = 0;
  for (uint32_t b = b0 + warp_id(); b < b1; b += blockDim.x >> 5) {
    const uint32_t w = intcomp::enc_block_size_warp(v, b);
    if (lane_id() == 0) {
      table[b] = w;
      acc += w;
    }
  }
  uint64_t tot;
  block_exclusive_scan(acc, ws, tot);
  if (threadIdx.x == 0) part[(uint64_t)blockIdx.x * ymax + y] = tot;
}

// one warp per huge list: 3 header words + the slices' partial sums + the var-byte tail
__global__ void __launch_bounds__(kCodecThreads)
k_enc_size_huge_fin(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off,
                    uint64_t nlists, uint64_t* __restrict__ sizes,
                    const uint64_t* __restrict__ part, uint32_t ymax,
                    const uint32_t* __restrict__ worklist, const uint32_t* __restrict__ n_work) {
  pdl_enter();
  const uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (h >= n_work[1]) return;
  const uint32_t i = worklist[nlists - 1 - h];
  const uint32_t* v = in + off[i];
  const uint32_t n = (uint32_t)(off[i + 1] - off[i]);
  const uint32_t nb = n >> 7, tail = n & 127u;
  const uint32_t Y = huge_slices(nb, ymax);
  uint64_t acc = 0;
  for (uint32_t y = lane_id(); y < Y; y += 32) acc += part[(uint64_t)h * ymax + y];
  uint32_t bytes = 0;
  for (uint32_t t = lane_id(); t < tail; t += 32) {
    const uint32_t idx = nb * 128 + t;
    bytes += intcomp::vbyte_len(intcomp::zigzag(v[idx], t ? v[idx - 1] : 0u));
  }
  bytes = warp_sum(bytes);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane_id() == 0) sizes[i] = 3 + acc + (tail ? 1 + (bytes + 3) / 4 : 0);
}

__global__ void __launch_bounds__(1024)
k_enc_emit_huge(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off, uint64_t nlists,
                const uint64_t* __restrict__ woff, uint32_t* __restrict__ words,
                uint32_t* __restrict__ tables, const uint64_t* __restrict__ part, uint32_t ymax,
                const uint32_t* __restrict__ worklist, const uint32_t* __restrict__ n_work) {
  pdl_enter();
  __shared__ uint64_t ws[1024 / 32 + 2];
  __shared__ uint32_t stage[32 * intcomp::kStageWords];
  if (blockIdx.x >= n_work[1]) return;
  const uint32_t i = worklist[nlists - 1 - blockIdx.x];
  const uint32_t* v = in + off[i];
  const uint32_t n = (uint32_t)(off[i + 1] - off[i]);
  const uint32_t nb = n >> 7, tail = n & 127u;
  const uint32_t Y = huge_slices(nb, ymax), y = blockIdx.y;
  if (y >= Y) return;
  const uint32_t b0 = (uint32_t)((uint64_t)nb * y / Y), b1 = (uint32_t)((uint64_t)nb * (y + 1) / Y);
  // the block table of list i lives at (off[i] - off[0]) / 128 of the shared table scratch
  uint32_t* table = tables + ((off[i] - off[0]) >> 7);
  uint32_t* dst = words + woff[i];
  // words of the slices before this one, and of all slices (ymax <= 1024 = the CTA)
  const uint64_t mine = threadIdx.x < Y ? part[(uint64_t)blockIdx.x * ymax + threadIdx.x] : (uint64_t)0;
  uint64_t total, before;
  block_exclusive_scan<uint64_t>(mine, ws, total);
  __syncthreads();
  block_exclusive_scan<uint64_t>(threadIdx.x < y ? mine : (uint64_t)0, ws, before);
  __syncthreads();
  uint64_t run = before;
  for (uint32_t base = b0; base < b1; base += blockDim.x) {
    const uint32_t b = base + threadIdx.x;
    const uint64_t x = b < b1 ? table[b] : 0u;
    uint64_t tot;
    const uint64_t ex = block_exclusive_scan(x, ws, tot);
    if (b < b1) table[b] = (uint32_t)(run + ex);
    run += tot;
  }
  __syncthreads();
  uint32_t* my_stage = stage + warp_id() * intcomp::kStageWords;
  for (uint32_t b = b0 + warp_id(); b < b1; b += blockDim.x >> 5)
    intcomp::enc_block_emit_warp(v, b, dst + 3 + table[b], my_stage);
  if (y == 0 && threadIdx.x == 0) {
    dst[0] = nb * 128;
    dst[1] = 3 + (uint32_t)total;
    dst[2] = v[0];
  }
  if (y == Y - 1 && tail && warp_id() == 0)
    intcomp::enc_tail_warp(v, nb, tail, dst + 3 + (uint32_t)total, my_stage);
}

__global__ void __launch_bounds__(kCodecThreads)
k_enc_size_long(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off,
                uint64_t* __restrict__ sizes, const uint32_t* __restrict__ worklist,
                const uint32_t* __restrict__ n_work) {
  pdl_enter();
  const uint32_t nw = *n_work;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nw; t += warps) {
    uint32_t i = worklist[t];
    uint64_t a = off[i];
    uint32_t w = intcomp::enc_size_warp(in + a, (uint32_t)(off[i + 1] - a));
    if (lane_id() == 0) sizes[i] = w;
  }
}

__global__ void __launch_bounds__(kCodecThreads)
k_enc_emit_short(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off, uint64_t nlists,
                 const uint64_t* __restrict__ woff, uint32_t* __restrict__ words) {
  pdl_enter();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlists) return;
  uint64_t a = off[i], n = off[i + 1] - a;
  if (n >= 128) return;
  intcomp::enc_emit_thread_small(in + a, (uint32_t)n, words + woff[i]);
}

__global__ void __launch_bounds__(kCodecThreads)
k_enc_emit_long(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off,
                const uint64_t* __restrict__ woff, uint32_t* __restrict__ words,
                const uint32_t* __restrict__ worklist, const uint32_t* __restrict__ n_work) {
  pdl_enter();
  __shared__ uint32_t stage[kCodecThreads / 32][intcomp::kStageWords];
  const uint32_t nw = *n_work;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nw; t += warps) {
    uint32_t i = worklist[t];
    uint64_t a = off[i];
    intcomp::enc_emit_warp(in + a, (uint32_t)(off[i + 1] - a), words + woff[i], stage[warp_id()]);
  }
}

int intcomp_encode_dev(const uint32_t* d_in, const uint64_t* d_off, uint64_t nlists,
                       uint64_t nvals_hint, DevBuf<uint32_t>& words, DevBuf<uint64_t>& woff,
                       uint64_t* total_words, cudaStream_t s) {
  ProfScope scope("k3a_encode", s);
  if (nlists >= (1ull << 32)) {
    set_last_error("more than 2^32 lists in one encode batch");
    return II2_ERR_UNSUPPORTED;
  }
  II2_TRY(woff.alloc_scratch(nlists + 1, s));
  DevBuf<uint32_t> n_work, worklist;
  DevBuf<uint64_t> d_total;
  II2_TRY(n_work.alloc_scratch(2, s));
  II2_TRY(d_total.alloc_scratch(1, s));
  II2_TRY(worklist.alloc_scratch(nlists ? nlists : 1, s));
  II2_CUDA_TRY(cudaMemsetAsync(n_work.p, 0, 2 * sizeof(uint32_t), s));
  II2_CUDA_TRY(cudaMemsetAsync(woff.p + nlists, 0, sizeof(uint64_t), s));
  // at most n_vals / kHugeList lists can be huge; n_vals is only known on the device here, so
  // the launch covers min(nlists, that bound from the caller) CTAs that exit when out of work
  const unsigned huge_grid = (unsigned)std::min<uint64_t>(nlists, nvals_hint / kHugeList + 1);
  // slices per huge list (grid.y): what the longest possible list could use, within 4096 CTAs
  // for the launch — the number of huge lists is only known on the device, CTAs past it or past
  // a list's own slice count exit at once, and 64 K of those were measured at +0.2 ms per batch
  const unsigned ymax = (unsigned)std::min<uint64_t>(
      1024, std::max<uint64_t>(1, std::min<uint64_t>((nvals_hint >> 7) / kSliceBlocks,
                                                     4096 / std::max(1u, huge_grid))));
  const dim3 huge_dim(huge_grid, ymax);
  DevBuf<uint32_t> tables;
  DevBuf<uint64_t> part;
  II2_TRY(tables.alloc_scratch((nvals_hint >> 7) + nlists + 1, s));
  II2_TRY(part.alloc_scratch((size_t)huge_grid * ymax + 1, s));
  if (nlists) {
    II2_LAUNCH_CHAIN(k_enc_size_short, div_up(nlists, kCodecThreads), kCodecThreads, 0, s, d_in, d_off, nlists, woff.p, worklist.p, n_work.p);
    II2_LAUNCH_CHAIN(k_enc_size_long, kNumSMs * 4, kCodecThreads, 0, s, d_in, d_off, woff.p, worklist.p, n_work.p);
    II2_LAUNCH_CHAIN(k_enc_size_huge, huge_dim, 1024, 0, s, d_in, d_off, nlists, tables.p, part.p, ymax, worklist.p, n_work.p);
    II2_LAUNCH_CHAIN(k_enc_size_huge_fin, div_up((uint64_t)huge_grid * 32, kCodecThreads), kCodecThreads, 0, s, d_in, d_off, nlists, woff.p, part.p, ymax, worklist.p, n_work.p);
  }
  II2_TRY(exclusive_scan_u64(woff.p, nlists + 1, d_total.p, s));
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 26, d_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  const uint64_t total = pinned_scratch()[26];
  II2_TRY(words.alloc_scratch(total, s));
  if (nlists) {
    II2_LAUNCH_CHAIN(k_enc_emit_short, div_up(nlists, kCodecThreads), kCodecThreads, 0, s, d_in, d_off, nlists, woff.p, words.p);
    II2_LAUNCH_CHAIN(k_enc_emit_long, kNumSMs * 4, kCodecThreads, 0, s, d_in, d_off, woff.p, words.p, worklist.p, n_work.p);
    II2_LAUNCH_CHAIN(k_enc_emit_huge, huge_dim, 1024, 0, s, d_in, d_off, nlists, woff.p, words.p, tables.p, part.p, ymax, worklist.p, n_work.p);
  }
  if (total_words) *total_words = total;
  return II2_OK;
}

}  // namespace ii2

// ---------------------------------------------------------------- C-ABI
using namespace ii2;

extern "C" {

int ii2_intcomp_encode_u32(const uint32_t* in, const uint64_t* off, uint64_t nlists,
                           uint32_t** words_out, uint64_t** word_off_out) {
  if (!words_out || !word_off_out || (nlists && !off)) return II2_ERR_INVALID;
  *words_out = nullptr;
  *word_off_out = nullptr;
  II2_TRY(ctx_require());
  cudaStream_t s = cur_stream();
  uint64_t first = nlists ? off[0] : 0, nvals = nlists ? off[nlists] - off[0] : 0;
  DevBuf<uint32_t> d_in, d_words;
  DevBuf<uint64_t> d_off, d_woff;
  II2_TRY(d_in.alloc_scratch(nvals, s));
  II2_TRY(d_off.alloc_scratch(nlists + 1, s));
  if (nvals)
    II2_CUDA_TRY(cudaMemcpyAsync(d_in.p, in + first, nvals * 4, cudaMemcpyHostToDevice, s));
  if (nlists) {
    II2_CUDA_TRY(cudaMemcpyAsync(d_off.p, off, (nlists + 1) * 8, cudaMemcpyHostToDevice, s));
  } else {
    II2_CUDA_TRY(cudaMemsetAsync(d_off.p, 0, 8, s));
  }
  // offsets are relative to `in`; the device copy starts at off[0]
  const uint32_t* d_base = d_in.p - first;
  uint64_t total = 0;
  II2_TRY(intcomp_encode_dev(d_base, d_off.p, nlists, nvals, d_words, d_woff, &total, s));
  uint32_t* hw = static_cast<uint32_t*>(pinned_alloc(total * 4 + 4));
  uint64_t* ho = static_cast<uint64_t*>(pinned_alloc((nlists + 1) * 8));
  if (!hw || !ho) {
    pinned_free(hw);
    pinned_free(ho);
    return II2_ERR_NOMEM;
  }
  if (total) II2_CUDA_TRY(cudaMemcpyAsync(hw, d_words.p, total * 4, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaMemcpyAsync(ho, d_woff.p, (nlists + 1) * 8, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  arena_reset(s);
  *words_out = hw;
  *word_off_out = ho;
  return II2_OK;
}

int ii2_intcomp_decode_u32(const uint32_t* words, const uint64_t* word_off, uint64_t nlists,
                           uint32_t** out, uint64_t** out_off) {
  if (!out || !out_off || (nlists && !word_off)) return II2_ERR_INVALID;
  *out = nullptr;
  *out_off = nullptr;
  II2_TRY(ctx_require());
  cudaStream_t s = cur_stream();
  uint64_t first = nlists ? word_off[0] : 0, nw = nlists ? word_off[nlists] - word_off[0] : 0;
  DevBuf<uint32_t> d_words, d_out;
  DevBuf<uint64_t> d_woff, d_ooff;
  II2_TRY(d_words.alloc_scratch(nw, s, 16));
  II2_TRY(d_woff.alloc_scratch(nlists + 1, s));
  if (nw) II2_CUDA_TRY(cudaMemcpyAsync(d_words.p, words + first, nw * 4, cudaMemcpyHostToDevice, s));
  if (nlists) {
    II2_CUDA_TRY(cudaMemcpyAsync(d_woff.p, word_off, (nlists + 1) * 8, cudaMemcpyHostToDevice, s));
  } else {
    II2_CUDA_TRY(cudaMemsetAsync(d_woff.p, 0, 8, s));
  }
  uint64_t total = 0;
  II2_TRY(intcomp_decode_dev(d_words.p - first, d_woff.p, nlists, d_out, d_ooff, &total, s, true));
  uint32_t* hv = static_cast<uint32_t*>(pinned_alloc(total * 4 + 4));
  uint64_t* ho = static_cast<uint64_t*>(pinned_alloc((nlists + 1) * 8));
  if (!hv || !ho) {
    pinned_free(hv);
    pinned_free(ho);
    return II2_ERR_NOMEM;
  }
  if (total) II2_CUDA_TRY(cudaMemcpyAsync(hv, d_out.p, total * 4, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaMemcpyAsync(ho, d_ooff.p, (nlists + 1) * 8, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  arena_reset(s);
  *out = hv;
  *out_off = ho;
  return II2_OK;
}

}  // extern "C"
