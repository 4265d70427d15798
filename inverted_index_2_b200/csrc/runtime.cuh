// runtime.cuh — process context, per-thread streams, stream-ordered device buffers,
// pinned host pool.  One process drives one GPU (torch.distributed-style launch).
#pragma once
#include <vector>

#include "common.cuh"

namespace ii2 {

bool ctx_ready();
int ctx_require();            // II2_OK or II2_ERR_NO_DEVICE (sets last error)
cudaStream_t cur_stream();    // this thread's stream (caller-provided or library-owned)

// Stream-ordered device buffer (cudaMallocAsync on the device's default pool).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; s = o.s;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  // `pad_bytes` extra readable bytes past the end (over-reading vector loads).
  int alloc(size_t count, cudaStream_t stream, size_t pad_bytes = 0) {
    release();
    s = stream;
    n = count;
    size_t bytes = count * sizeof(T) + pad_bytes;
    if (bytes == 0) bytes = 16;
    void* q = nullptr;
    II2_CUDA_TRY(cudaMallocAsync(&q, bytes, stream));
    p = static_cast<T*>(q);
    return II2_OK;
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  T* take() { T* q = p; p = nullptr; n = 0; return q; }
};

// Instrumentation (ii2_prof_enable / ii2_prof_read): CUDA events around a kernel phase on the
// launching stream.  No-ops unless enabled.
struct ProfScope {
  cudaStream_t s;
  int slot;
  ProfScope(const char* name, cudaStream_t stream);
  void end();
  ~ProfScope() { end(); }
};

// Pinned host memory with a size-class cache (cudaHostAlloc is far too slow per call).
void* pinned_alloc(size_t bytes);
void pinned_free(void* p);  // also accepts nullptr

// In-place exclusive scan of d[0..n) (u64); total written to *d_total (device) if non-null.
int exclusive_scan_u64(uint64_t* d, uint64_t n, uint64_t* d_total, cudaStream_t s);
// m short arrays (each n <= a few thousand) laid out back to back, scanned independently by
// one CTA in one launch; totals[j] (device) receives each array's sum.
// (in == out is allowed.)
int exclusive_scan_multi_u64(const uint64_t* in, uint64_t* out, uint64_t n, int m,
                             uint64_t* d_totals, cudaStream_t s);

}  // namespace ii2
