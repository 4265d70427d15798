// k3a_intcomp.cu — batched posting codec kernels (K3a): many independent lists per launch.
// Replaces one intcomp.CompressUint32 call per term (file/writer.go:49) and one
// intcomp.UncompressUint32 per term (file/reader.go:100).
//   size pass -> exclusive scan (so `_val` offsets equal the reference's running
//   valuesOffset, file/writer.go:56) -> emit pass.
// Lists below one 128-block are handled one per thread; longer lists one per warp
// (32-lane groups match the codec's 32-value groups).
#include <algorithm>

#include "codec.cuh"
#include "intcomp.cuh"
#include "runtime.cuh"

namespace ii2 {

constexpr int kCodecThreads = 256;

// ---------------------------------------------------------------- decode
__global__ void __launch_bounds__(kCodecThreads)
k_dec_count(const uint32_t* __restrict__ words, const uint64_t* __restrict__ woff, uint64_t nlists,
            uint64_t* __restrict__ counts, int* __restrict__ err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlists) return;
  uint64_t a = woff[i], b = woff[i + 1];
  long long c = b >= a ? intcomp::dec_count(words + a, b - a) : -1;
  if (c < 0) {
    atomicExch(err, 1);
    c = 0;
  }
  counts[i] = (uint64_t)c;
}

// FST outputs (byte offsets) + file size -> word offsets [n+1] (file/reader.go:52,64)
__global__ void __launch_bounds__(kCodecThreads)
k_valoff_to_woff(const uint64_t* __restrict__ val_off, uint64_t n, uint64_t val_size,
                 uint64_t* __restrict__ woff, int* __restrict__ err) {
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
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlists) return;
  uint64_t a = woff[i], b = woff[i + 1];
  if (b <= a) return;  // empty run -> empty list (file/writer_test.go:15)
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
  DevBuf<uint32_t> n_work, worklist;
  DevBuf<uint64_t> d_total;
  II2_TRY(err.alloc_scratch(1, s));
  II2_TRY(n_work.alloc_scratch(1, s));
  II2_TRY(d_total.alloc_scratch(1, s));
  II2_TRY(worklist.alloc_scratch(nlists ? nlists : 1, s));
  II2_CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int), s));
  II2_CUDA_TRY(cudaMemsetAsync(n_work.p, 0, sizeof(uint32_t), s));
  II2_CUDA_TRY(cudaMemsetAsync(out_off.p + nlists, 0, sizeof(uint64_t), s));
  if (nlists) {
    k_dec_count<<<div_up(nlists, kCodecThreads), kCodecThreads, 0, s>>>(d_words, d_woff, nlists,
                                                                        out_off.p, err.p);
    II2_LAUNCHED();
  }
  II2_TRY(exclusive_scan_u64(out_off.p, nlists + 1, d_total.p, s));
  uint64_t total = 0;
  int herr = 0;
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 20, d_total.p, sizeof(total), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaMemcpyAsync(pinned_scratch() + 21, err.p, sizeof(herr), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  total = pinned_scratch()[20];
  herr = *reinterpret_cast<const int*>(pinned_scratch() + 21);
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
    k_dec_short<<<div_up(nlists, kCodecThreads), kCodecThreads, 0, s>>>(
        d_words, d_woff, nlists, out_off.p, out.p, worklist.p, n_work.p, err.p);
    II2_LAUNCHED();
    k_dec_long<<<kNumSMs * 4, kCodecThreads, 0, s>>>(d_words, d_woff, out_off.p, out.p, worklist.p,
                                                     n_work.p, err.p);
    II2_LAUNCHED();
  }
  II2_CUDA_TRY(cudaMemcpyAsync(&herr, err.p, sizeof(herr), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
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
  k_valoff_to_woff<<<div_up(n + 1, kCodecThreads), kCodecThreads, 0, s>>>(d_val_off, n, val_size,
                                                                          woff.p, err.p);
  II2_LAUNCHED();
  int herr = 0;
  II2_CUDA_TRY(cudaMemcpyAsync(&herr, err.p, sizeof(herr), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  if (herr) {
    set_last_error("_val offsets are not 4-byte aligned or exceed the file size");
    return II2_ERR_CORRUPT;
  }
  return II2_OK;
}

// ---------------------------------------------------------------- encode
constexpr uint64_t kHugeList = 8192;  // values from which a whole CTA encodes one list

// n_work[0] counts the warp-per-list entries (filled from the front of `worklist`), n_work[1]
// the CTA-per-list entries (filled from the back)
__global__ void __launch_bounds__(kCodecThreads)
k_enc_size_short(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off, uint64_t nlists,
                 uint64_t* __restrict__ sizes, uint32_t* __restrict__ worklist,
                 uint32_t* __restrict__ n_work) {
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

// one CTA per huge list: Σ block sizes + tail
__global__ void __launch_bounds__(1024)
k_enc_size_huge(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off, uint64_t nlists,
                uint64_t* __restrict__ sizes, const uint32_t* __restrict__ worklist,
                const uint32_t* __restrict__ n_work) {
  __shared__ uint64_t ws[1024 / 32 + 2];
  if (blockIdx.x >= n_work[1]) return;
  const uint32_t i = worklist[nlists - 1 - blockIdx.x];
  const uint32_t* v = in + off[i];
  const uint32_t n = (uint32_t)(off[i + 1] - off[i]);
  const uint32_t nb = n >> 7, tail = n & 127u;
  uint64_t acc = 0;
  for (uint32_t b = warp_id(); b < nb; b += blockDim.x >> 5) {
    const uint32_t w = intcomp::enc_block_size_warp(v, b);
    if (lane_id() == 0) acc += w;
  }
  if (tail && threadIdx.x < 32) {
    uint32_t bytes = 0;
    for (uint32_t t = threadIdx.x; t < tail; t += 32) {
      const uint32_t idx = nb * 128 + t;
      bytes += intcomp::vbyte_len(intcomp::zigzag(v[idx], t ? v[idx - 1] : 0u));
    }
    bytes = warp_sum(bytes);
    if (threadIdx.x == 0) acc += 1 + (bytes + 3) / 4;
  }
  uint64_t tot;
  block_exclusive_scan(acc, ws, tot);
  if (threadIdx.x == 0) sizes[i] = 3 + tot;
}

__global__ void __launch_bounds__(1024)
k_enc_emit_huge(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off, uint64_t nlists,
                const uint64_t* __restrict__ woff, uint32_t* __restrict__ words,
                uint32_t* __restrict__ tables, const uint32_t* __restrict__ worklist,
                const uint32_t* __restrict__ n_work) {
  __shared__ uint64_t ws[1024 / 32 + 2];
  __shared__ uint32_t stage[32 * intcomp::kStageWords];
  if (blockIdx.x >= n_work[1]) return;
  const uint32_t i = worklist[nlists - 1 - blockIdx.x];
  // the block table of list i lives at (off[i] - off[0]) / 128 of the shared table scratch
  intcomp::enc_emit_cta(in + off[i], (uint32_t)(off[i + 1] - off[i]), words + woff[i],
                        tables + ((off[i] - off[0]) >> 7), stage, ws);
}

__global__ void __launch_bounds__(kCodecThreads)
k_enc_size_long(const uint32_t* __restrict__ in, const uint64_t* __restrict__ off,
                uint64_t* __restrict__ sizes, const uint32_t* __restrict__ worklist,
                const uint32_t* __restrict__ n_work) {
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
  DevBuf<uint32_t> tables;
  II2_TRY(tables.alloc_scratch((nvals_hint >> 7) + nlists + 1, s));
  if (nlists) {
    k_enc_size_short<<<div_up(nlists, kCodecThreads), kCodecThreads, 0, s>>>(
        d_in, d_off, nlists, woff.p, worklist.p, n_work.p);
    II2_LAUNCHED();
    k_enc_size_long<<<kNumSMs * 4, kCodecThreads, 0, s>>>(d_in, d_off, woff.p, worklist.p,
                                                          n_work.p);
    II2_LAUNCHED();
    k_enc_size_huge<<<huge_grid, 1024, 0, s>>>(d_in, d_off, nlists, woff.p, worklist.p, n_work.p);
    II2_LAUNCHED();
  }
  II2_TRY(exclusive_scan_u64(woff.p, nlists + 1, d_total.p, s));
  uint64_t total = 0;
  II2_CUDA_TRY(cudaMemcpyAsync(&total, d_total.p, sizeof(total), cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  II2_TRY(words.alloc_scratch(total, s));
  if (nlists) {
    k_enc_emit_short<<<div_up(nlists, kCodecThreads), kCodecThreads, 0, s>>>(d_in, d_off, nlists,
                                                                             woff.p, words.p);
    II2_LAUNCHED();
    k_enc_emit_long<<<kNumSMs * 4, kCodecThreads, 0, s>>>(d_in, d_off, woff.p, words.p,
                                                          worklist.p, n_work.p);
    II2_LAUNCHED();
    k_enc_emit_huge<<<huge_grid, 1024, 0, s>>>(d_in, d_off, nlists, woff.p, words.p, tables.p,
                                               worklist.p, n_work.p);
    II2_LAUNCHED();
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
