// api.cu — C-ABI entry points for compaction and term-range reads (include/ii2.h):
// resident segments / removed sets / results, the device pipelines, and the host-buffer
// convenience calls that stage through them.
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "codec.cuh"
#include "union.cuh"

using namespace ii2;

// ------------------------------------------------------------------ opaque handle types
struct ii2_seg {
  uint32_t n_terms = 0;
  uint64_t n_post = 0;
  uint64_t term_bytes_len = 0;
  DevBuf<uint8_t> tb;
  DevBuf<uint32_t> toff;
  DevBuf<uint32_t> post;
  DevBuf<uint64_t> poff;
};

struct ii2_removed {
  uint64_t n = 0;
  DevBuf<uint32_t> sorted;
  DevBuf<uint32_t> bitmap;
  uint64_t bitmap_bits = 0;
  RemovedSet set() const {
    RemovedSet r;
    r.sorted = sorted.p;
    r.n = n;
    r.bitmap = bitmap_bits ? bitmap.p : nullptr;
    r.bitmap_bits = bitmap_bits;
    return r;
  }
};

struct ii2_result {
  EmitOut out;
  uint64_t T = 0, TB = 0, P = 0, E = 0;
  uint64_t postings_in = 0, terms_merged = 0;
  bool has_dec = false, has_enc = false;
  bool has_minmax = false;
  std::string min_term, max_term;
};

namespace {

// ------------------------------------------------------------------ small kernels
__global__ void __launch_bounds__(256)
k_direct_to_lists(const uint64_t* __restrict__ val_off, uint32_t n, uint32_t* __restrict__ post,
                  uint64_t* __restrict__ poff) {
  // direct mode: the FST output is the single posting, truncated to uint32 (file/reader.go:75)
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    post[i] = (uint32_t)val_off[i];
    poff[i] = i;
  }
  if (i == n) poff[n] = n;
}

// max term length, and whether offsets are monotone
__global__ void __launch_bounds__(256)
k_seg_check(const uint32_t* __restrict__ toff, uint32_t n, const uint64_t* __restrict__ poff,
            uint32_t* __restrict__ stats /* [0]=max term len, [1]=bad flag */) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a = toff[i], b = toff[i + 1];
  if (b < a) atomicExch(&stats[1], 1u);
  else atomicMax(&stats[0], b - a);
  if (poff && poff[i + 1] < poff[i]) atomicExch(&stats[1], 1u);
}

__global__ void __launch_bounds__(256)
k_bitmap_set(const uint32_t* __restrict__ sorted, uint64_t n, uint32_t* __restrict__ bitmap) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicOr(&bitmap[sorted[i] >> 5], 1u << (sorted[i] & 31u));
}

// Range windows (K4): lo = first term >= min (vellum Iterator(min) seek, file/reader.go:147),
// hi = first term > max (inclusive right bound, :54-58 and :151-155); then bases.  One CTA.
__global__ void __launch_bounds__(1024)
k4_windows(SegDesc* __restrict__ segs, int k, const uint8_t* __restrict__ bounds, uint32_t minlen,
           int has_min, uint32_t maxlen, int has_max, uint32_t* __restrict__ n_total) {
  __shared__ uint32_t ws[1024 / 32 + 2];
  const int s = threadIdx.x;
  uint32_t lo = 0, hi = 0;
  if (s < k) {
    SegDesc sd = segs[s];
    lo = has_min ? seg_lower_bound(sd, 0, sd.n, bounds, minlen) : 0u;
    hi = has_max ? seg_upper_bound(sd, lo, sd.n, bounds + minlen, maxlen) : sd.n;
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan(hi - lo, ws, tot);
  if (s < k) {
    segs[s].lo = lo;
    segs[s].hi = hi;
    segs[s].base = ex;
  }
  if (s == 0) *n_total = tot;
}

// ------------------------------------------------------------------ the device pipeline
// segs: resident segments; [min,max] optional; rem optional.
int run_pipeline_impl(ii2_seg* const* segs, int nseg, const uint8_t* min, size_t minlen,
                      bool has_min, const uint8_t* max, size_t maxlen, bool has_max,
                      const ii2_removed* rem, bool want_dec, bool want_enc, bool want_minmax,
                      bool keep_empty, ii2_result** res_out, cudaStream_t s) {
  ProfScope pipe_scope("pipeline_total", s);
  std::unique_ptr<ii2_result> res(new ii2_result());
  res->has_dec = want_dec;
  res->has_enc = want_enc;
  if (nseg < 0 || (nseg > 0 && !segs)) return II2_ERR_INVALID;
  if (nseg > kMaxSegs) {
    set_last_error("%d segments in one call (max %d per pass)", nseg, kMaxSegs);
    return II2_ERR_UNSUPPORTED;
  }
  // segment table, sample bases and range bounds travel through one pinned block so that no
  // pageable copy stalls the stream
  struct PinnedBlock {
    void* p = nullptr;
    ~PinnedBlock() { pinned_free(p); }
  } stage;
  const size_t nsegx = nseg ? nseg : 1;
  const size_t stage_bytes = sizeof(SegDesc) * nsegx + 8 * (nsegx + 1) + minlen + maxlen + 64;
  stage.p = pinned_alloc(stage_bytes);
  if (!stage.p) return II2_ERR_NOMEM;
  SegDesc* h = static_cast<SegDesc*>(stage.p);
  uint32_t* h_sbase = reinterpret_cast<uint32_t*>(h + nsegx);
  uint8_t* h_bounds = reinterpret_cast<uint8_t*>(h_sbase + 2 * (nsegx + 1));
  uint64_t n_total64 = 0, n_in = 0;
  for (int i = 0; i < nseg; i++) {
    const ii2_seg* g = segs[i];
    if (!g) return II2_ERR_INVALID;
    h[i].tb = g->tb.p;
    h[i].toff = g->toff.p;
    h[i].post = g->post.p;
    h[i].poff = g->poff.p;
    h[i].n = g->n_terms;
    h[i].lo = 0;
    h[i].hi = g->n_terms;
    h[i].base = (uint32_t)n_total64;
    n_total64 += g->n_terms;
    n_in += g->n_post;
  }
  if (n_total64 >= (1ull << 32)) {
    set_last_error("more than 2^32-1 term instances in one call");
    return II2_ERR_UNSUPPORTED;
  }
  DevBuf<SegDesc> d_segs;
  II2_TRY(d_segs.alloc_scratch(nsegx, s));
  if (nseg)
    II2_CUDA_TRY(cudaMemcpyAsync(d_segs.p, h, sizeof(SegDesc) * nseg, cudaMemcpyHostToDevice, s));
  uint32_t n_total = (uint32_t)n_total64;
  const bool ranged = has_min || has_max;
  if (ranged && nseg) {
    DevBuf<uint8_t> d_bounds;
    DevBuf<uint32_t> d_nt;
    II2_TRY(d_bounds.alloc_scratch(minlen + maxlen + 8, s));
    II2_TRY(d_nt.alloc_scratch(1, s));
    if (has_min && minlen) memcpy(h_bounds, min, minlen);
    if (has_max && maxlen) memcpy(h_bounds + minlen, max, maxlen);
    if (minlen + maxlen)
      II2_CUDA_TRY(cudaMemcpyAsync(d_bounds.p, h_bounds, minlen + maxlen, cudaMemcpyHostToDevice, s));
    k4_windows<<<1, 1024, 0, s>>>(d_segs.p, nseg, d_bounds.p, (uint32_t)minlen, has_min ? 1 : 0,
                                  (uint32_t)maxlen, has_max ? 1 : 0, d_nt.p);
    II2_LAUNCHED();
    // the windows come back: the planner spreads its samples over them
    II2_CUDA_TRY(cudaMemcpyAsync(h, d_segs.p, sizeof(SegDesc) * nseg, cudaMemcpyDeviceToHost, s));
    II2_CUDA_TRY(cudaStreamSynchronize(s));
    n_total64 = 0;
    for (int i = 0; i < nseg; i++) n_total64 += h[i].hi - h[i].lo;
    n_total = (uint32_t)n_total64;
  }

  EmitOut& out = res->out;
  if (n_total == 0) {  // nothing in range / no terms at all: empty result
    II2_TRY(out.term_bytes.alloc(0, s, 32));
    II2_TRY(out.term_off.alloc(1, s));
    II2_CUDA_TRY(cudaMemsetAsync(out.term_off.p, 0, 4, s));
    if (want_dec) {
      II2_TRY(out.post.alloc(0, s));
      II2_TRY(out.post_off.alloc(1, s));
      II2_CUDA_TRY(cudaMemsetAsync(out.post_off.p, 0, 8, s));
    }
    if (want_enc) {
      II2_TRY(out.val_words.alloc(0, s));
      II2_TRY(out.val_off.alloc(0, s));
    }
    II2_CUDA_TRY(cudaStreamSynchronize(s));
    *res_out = res.release();
    return II2_OK;
  }

  MergePlan plan;
  plan.k = nseg;
  plan.n_total = n_total;
  plan.segs = d_segs.p;
  II2_TRY(k1_build_plan(plan, h, h_sbase, s));
  if (ranged) {  // Σ input postings inside the windows sizes the union buffers
    uint64_t h_plan_tot[2] = {0, 0};
    II2_CUDA_TRY(cudaMemcpyAsync(h_plan_tot, plan.totals.p, 16, cudaMemcpyDeviceToHost, s));
    II2_CUDA_TRY(cudaStreamSynchronize(s));
    n_in = h_plan_tot[1];
  }

  RemovedSet rs;
  if (rem) {
    rs = rem->set();
  } else {
    rs.sorted = nullptr;
    rs.n = 0;
    rs.bitmap = nullptr;
    rs.bitmap_bits = 0;
  }
  UnionOut u;
  II2_TRY(k12_union(plan, rs, want_dec, want_enc, keep_empty, n_in, u, s));
  res->T = u.h_totals[0];
  res->TB = u.h_totals[1];
  res->P = u.h_totals[2];
  res->E = u.h_totals[3];
  res->postings_in = n_in;
  res->terms_merged = u.terms_merged;
  DevBuf<uint8_t> d_mm;
  if (want_minmax) {
    II2_TRY(d_mm.alloc_scratch(8 + 2 * 65536, s));
    II2_TRY(k6_minmax(plan, u, d_mm.p, s));
  }
  II2_TRY(k6_emit(plan, u, out, s));

  uint8_t* mm = nullptr;
  if (want_minmax) {
    mm = static_cast<uint8_t*>(pinned_alloc(8 + 2 * 65536));
    if (!mm) return II2_ERR_NOMEM;
    cudaError_t e = cudaMemcpyAsync(mm, d_mm.p, 4096, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    uint32_t ln[2] = {0, 0};
    if (e == cudaSuccess) {
      memcpy(ln, mm, 8);
      if (8 + (size_t)ln[0] + ln[1] > 4096) {  // long terms: fetch the rest
        e = cudaMemcpyAsync(mm, d_mm.p, 8 + (size_t)ln[0] + ln[1], cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      }
    }
    if (e != cudaSuccess) {
      pinned_free(mm);
      set_last_error("min/max copy: %s", cudaGetErrorString(e));
      return II2_ERR_CUDA;
    }
    res->min_term.assign(reinterpret_cast<const char*>(mm) + 8, ln[0]);
    res->max_term.assign(reinterpret_cast<const char*>(mm) + 8 + ln[0], ln[1]);
    res->has_minmax = u.terms_merged > 0;
    pinned_free(mm);
  } else {
    II2_CUDA_TRY(cudaStreamSynchronize(s));
  }
  *res_out = res.release();
  return II2_OK;
}

// Intermediates live in the calling thread's scratch arena for the duration of the call.
int run_pipeline(ii2_seg* const* segs, int nseg, const uint8_t* min, size_t minlen, bool has_min,
                 const uint8_t* max, size_t maxlen, bool has_max, const ii2_removed* rem,
                 bool want_dec, bool want_enc, bool want_minmax, bool keep_empty,
                 ii2_result** res_out) {
  cudaStream_t s = cur_stream();
  const int rc = run_pipeline_impl(segs, nseg, min, minlen, has_min, max, maxlen, has_max, rem,
                                   want_dec, want_enc, want_minmax, keep_empty, res_out, s);
  if (rc != II2_OK) cudaStreamSynchronize(s);
  arena_reset(s);
  return rc;
}

template <typename T>
int h2d(DevBuf<T>& dst, const T* src, size_t n, cudaStream_t s, size_t pad = 0) {
  II2_TRY(dst.alloc(n, s, pad));
  if (n) II2_CUDA_TRY(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
  return II2_OK;
}

struct HostOwner {
  std::vector<void*> ptrs;
  ~HostOwner() {
    for (void* p : ptrs) pinned_free(p);
  }
  template <typename T>
  T* alloc(size_t n) {
    void* p = pinned_alloc(n * sizeof(T) + 8);
    if (p) ptrs.push_back(p);
    return static_cast<T*>(p);
  }
};

template <typename T>
int d2h(T** dst, const T* src, size_t n, HostOwner& own, cudaStream_t s) {
  T* p = own.alloc<T>(n);
  if (!p) return II2_ERR_NOMEM;
  if (n) II2_CUDA_TRY(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyDeviceToHost, s));
  *dst = p;
  return II2_OK;
}

}  // namespace

// ------------------------------------------------------------------ C-ABI
extern "C" {

int ii2_seg_upload(const ii2_seg_view* v, ii2_seg** seg_out) {
  if (!v || !seg_out) return II2_ERR_INVALID;
  *seg_out = nullptr;
  II2_TRY(ctx_require());
  cudaStream_t s = cur_stream();
  const uint64_t n = v->n_terms;
  if (n >= 0xFFFFFFFFull) {
    set_last_error("segment with %llu terms (max 2^32-2)", (unsigned long long)n);
    return II2_ERR_UNSUPPORTED;
  }
  if (n && (!v->term_off || (v->term_off[n] && !v->term_bytes))) return II2_ERR_INVALID;
  std::unique_ptr<ii2_seg> g(new ii2_seg());
  g->n_terms = (uint32_t)n;
  g->term_bytes_len = n ? v->term_off[n] : 0;
  II2_TRY(h2d(g->tb, v->term_bytes, (size_t)g->term_bytes_len, s, 32));
  if (n) {
    II2_TRY(h2d(g->toff, v->term_off, (size_t)n + 1, s));
  } else {
    II2_TRY(g->toff.alloc(1, s));
    II2_CUDA_TRY(cudaMemsetAsync(g->toff.p, 0, 4, s));
  }
  if (v->mode == II2_SEG_DECODED) {
    if (n && !v->post_off) return II2_ERR_INVALID;
    const uint64_t first = n ? v->post_off[0] : 0;
    g->n_post = n ? v->post_off[n] - first : 0;
    if (g->n_post && !v->post) return II2_ERR_INVALID;
    II2_TRY(h2d(g->post, v->post ? v->post + first : nullptr, (size_t)g->n_post, s, 16));
    if (n && first == 0) {
      II2_TRY(h2d(g->poff, v->post_off, (size_t)n + 1, s));
    } else {
      std::vector<uint64_t> rebased(n + 1, 0);
      for (uint64_t i = 0; i <= n && n; i++) rebased[i] = v->post_off[i] - first;
      II2_TRY(h2d(g->poff, rebased.data(), (size_t)n + 1, s));
      II2_CUDA_TRY(cudaStreamSynchronize(s));
    }
  } else if (v->mode == II2_SEG_DIRECT) {
    if (n && !v->val_off) return II2_ERR_INVALID;
    DevBuf<uint64_t> d_vo;
    II2_TRY(h2d(d_vo, v->val_off, (size_t)n, s));
    II2_TRY(g->post.alloc(n, s, 16));
    II2_TRY(g->poff.alloc(n + 1, s));
    k_direct_to_lists<<<div_up(n + 1, 256), 256, 0, s>>>(d_vo.p, (uint32_t)n, g->post.p, g->poff.p);
    II2_LAUNCHED();
    g->n_post = n;
    II2_CUDA_TRY(cudaStreamSynchronize(s));
  } else if (v->mode == II2_SEG_VAL) {
    if (n && !v->val_off) return II2_ERR_INVALID;
    if (v->val_size && !v->val_bytes) return II2_ERR_INVALID;
    if (v->val_size & 3) {
      set_last_error("_val size %llu is not a multiple of 4", (unsigned long long)v->val_size);
      return II2_ERR_CORRUPT;
    }
    DevBuf<uint64_t> d_vo, d_woff;
    DevBuf<uint32_t> d_words;
    II2_TRY(h2d(d_vo, v->val_off, (size_t)n, s));
    II2_TRY(d_words.alloc_scratch(v->val_size / 4, s, 16));
    if (v->val_size)
      II2_CUDA_TRY(cudaMemcpyAsync(d_words.p, v->val_bytes, v->val_size, cudaMemcpyHostToDevice, s));
    II2_TRY(val_offsets_to_word_offsets(d_vo.p, n, v->val_size, d_woff, s));
    uint64_t total = 0;
    II2_TRY(intcomp_decode_dev(d_words.p, d_woff.p, n, g->post, g->poff, &total, s, false));
    g->n_post = total;
  } else {
    set_last_error("unknown segment mode %d", v->mode);
    return II2_ERR_INVALID;
  }
  // sanity: term lengths must fit the tile kernel's 16-bit length field
  DevBuf<uint32_t> stats;
  II2_TRY(stats.alloc_scratch(2, s));
  II2_CUDA_TRY(cudaMemsetAsync(stats.p, 0, 8, s));
  if (n) {
    k_seg_check<<<div_up(n, 256), 256, 0, s>>>(g->toff.p, (uint32_t)n, g->poff.p, stats.p);
    II2_LAUNCHED();
  }
  uint32_t hs[2] = {0, 0};
  II2_CUDA_TRY(cudaMemcpyAsync(hs, stats.p, 8, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  if (hs[1]) {
    set_last_error("segment offsets are not monotone");
    return II2_ERR_INVALID;
  }
  if (hs[0] > 65535) {
    set_last_error("term of %u bytes (max 65535)", hs[0]);
    return II2_ERR_UNSUPPORTED;
  }
  arena_reset(s);
  *seg_out = g.release();
  return II2_OK;
}

void ii2_seg_release(ii2_seg* seg) { delete seg; }

int ii2_removed_upload(const uint32_t* removed_sorted, uint64_t nrem, ii2_removed** out) {
  if (!out || (nrem && !removed_sorted)) return II2_ERR_INVALID;
  *out = nullptr;
  II2_TRY(ctx_require());
  cudaStream_t s = cur_stream();
  std::unique_ptr<ii2_removed> r(new ii2_removed());
  if (nrem == 0) removed_sorted = nullptr;
  r->n = nrem;
  II2_TRY(h2d(r->sorted, removed_sorted, (size_t)nrem, s));
  if (nrem) {
    // membership bitmap when the id range is small enough to stay L2-friendly (<= 64 MiB)
    const uint64_t maxv = removed_sorted[nrem - 1];
    if (maxv < (1ull << 29) && nrem >= 64) {
      r->bitmap_bits = (maxv + 32) & ~31ull;
      II2_TRY(r->bitmap.alloc(r->bitmap_bits / 32, s));
      II2_CUDA_TRY(cudaMemsetAsync(r->bitmap.p, 0, r->bitmap_bits / 8, s));
      k_bitmap_set<<<div_up(nrem, 256), 256, 0, s>>>(r->sorted.p, nrem, r->bitmap.p);
      II2_LAUNCHED();
    }
  }
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  *out = r.release();
  return II2_OK;
}

void ii2_removed_release(ii2_removed* rem) { delete rem; }

int ii2_merge_dev(ii2_seg* const* segs, int nseg, const ii2_removed* rem, uint32_t flags,
                  ii2_result** res) {
  if (!res) return II2_ERR_INVALID;
  *res = nullptr;
  II2_TRY(ctx_require());
  if (flags == 0) flags = II2_RESULT_DECODED;
  // merge: full windows, min/max recorded, emptied terms dropped
  return run_pipeline(segs, nseg, nullptr, 0, false, nullptr, 0, false, rem,
                      (flags & II2_RESULT_DECODED) != 0, (flags & II2_RESULT_ENCODED) != 0, true,
                      false, res);
}

int ii2_read_range_dev(ii2_seg* const* segs, int nseg, const uint8_t* min, size_t minlen,
                       const uint8_t* max, size_t maxlen, const ii2_removed* rem,
                       ii2_result** res) {
  if (!res) return II2_ERR_INVALID;
  *res = nullptr;
  II2_TRY(ctx_require());
  // plain reads keep empty lists (file/writer_test.go:15 round trip); with a removed list the
  // merge-style filter drops emptied terms
  return run_pipeline(segs, nseg, min, minlen, min != nullptr, max, maxlen, max != nullptr, rem,
                      true, false, false, rem == nullptr, res);
}

int ii2_result_info_get(const ii2_result* r, ii2_result_info* info) {
  if (!r || !info) return II2_ERR_INVALID;
  info->terms_count = r->T;
  info->term_bytes = r->TB;
  info->postings_out = r->P;
  info->postings_in = r->postings_in;
  info->terms_merged = r->terms_merged;
  info->val_size = r->E * 4;
  info->d_term_bytes = r->out.term_bytes.p;
  info->d_term_off = r->out.term_off.p;
  info->d_post = r->out.post.p;
  info->d_post_off = r->out.post_off.p;
  info->d_val_bytes = r->out.val_words.p;
  info->d_val_off = r->out.val_off.p;
  return II2_OK;
}

int ii2_result_download_merge(const ii2_result* r, uint32_t flags, ii2_merge_out* o) {
  if (!r || !o) return II2_ERR_INVALID;
  memset(o, 0, sizeof(*o));
  II2_TRY(ctx_require());
  if (!r->has_enc) {
    set_last_error("result was produced without the encoder");
    return II2_ERR_INVALID;
  }
  cudaStream_t s = cur_stream();
  std::unique_ptr<HostOwner> own(new HostOwner());
  II2_TRY(d2h(&o->term_bytes, r->out.term_bytes.p, (size_t)r->TB, *own, s));
  II2_TRY(d2h(&o->term_off, r->out.term_off.p, (size_t)r->T + 1, *own, s));
  II2_TRY(d2h(&o->val_off, r->out.val_off.p, (size_t)r->T, *own, s));
  uint32_t* words = nullptr;
  II2_TRY(d2h(&words, r->out.val_words.p, (size_t)r->E, *own, s));
  o->val_bytes = reinterpret_cast<uint8_t*>(words);
  o->val_size = r->E * 4;
  if ((flags & II2_MERGE_WANT_DECODED) && r->has_dec) {
    II2_TRY(d2h(&o->post, r->out.post.p, (size_t)r->P, *own, s));
    II2_TRY(d2h(&o->post_off, r->out.post_off.p, (size_t)r->T + 1, *own, s));
  }
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  o->terms_count = r->T;
  o->has_minmax = r->has_minmax ? 1 : 0;
  if (r->has_minmax) {
    o->min_term = own->alloc<uint8_t>(r->min_term.size() + 1);
    o->max_term = own->alloc<uint8_t>(r->max_term.size() + 1);
    if (!o->min_term || !o->max_term) return II2_ERR_NOMEM;
    memcpy(o->min_term, r->min_term.data(), r->min_term.size());
    memcpy(o->max_term, r->max_term.data(), r->max_term.size());
    o->min_term_len = (uint32_t)r->min_term.size();
    o->max_term_len = (uint32_t)r->max_term.size();
  }
  o->terms_merged = r->terms_merged;
  o->postings_in = r->postings_in;
  o->postings_out = r->P;
  o->_owner = own.release();
  return II2_OK;
}

int ii2_result_download_read(const ii2_result* r, ii2_read_out* o) {
  if (!r || !o) return II2_ERR_INVALID;
  memset(o, 0, sizeof(*o));
  II2_TRY(ctx_require());
  if (!r->has_dec) return II2_ERR_INVALID;
  cudaStream_t s = cur_stream();
  std::unique_ptr<HostOwner> own(new HostOwner());
  II2_TRY(d2h(&o->term_bytes, r->out.term_bytes.p, (size_t)r->TB, *own, s));
  II2_TRY(d2h(&o->term_off, r->out.term_off.p, (size_t)r->T + 1, *own, s));
  II2_TRY(d2h(&o->post, r->out.post.p, (size_t)r->P, *own, s));
  II2_TRY(d2h(&o->post_off, r->out.post_off.p, (size_t)r->T + 1, *own, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  o->n_terms = r->T;
  o->_owner = own.release();
  return II2_OK;
}

int ii2_result_to_seg(ii2_result* r, ii2_seg** seg_out) {
  if (!r || !seg_out) return II2_ERR_INVALID;
  *seg_out = nullptr;
  if (!r->has_dec || !r->out.post_off.p) return II2_ERR_INVALID;
  if (r->T >= 0xFFFFFFFFull) return II2_ERR_UNSUPPORTED;
  std::unique_ptr<ii2_seg> g(new ii2_seg());
  g->n_terms = (uint32_t)r->T;
  g->n_post = r->P;
  g->term_bytes_len = r->TB;
  g->tb = std::move(r->out.term_bytes);
  g->toff = std::move(r->out.term_off);
  g->post = std::move(r->out.post);
  g->poff = std::move(r->out.post_off);
  r->has_dec = false;
  r->T = r->TB = r->P = 0;
  *seg_out = g.release();
  return II2_OK;
}

void ii2_result_release(ii2_result* r) { delete r; }

void ii2_merge_out_free(ii2_merge_out* o) {
  if (!o) return;
  delete static_cast<HostOwner*>(o->_owner);
  memset(o, 0, sizeof(*o));
}

void ii2_read_out_free(ii2_read_out* o) {
  if (!o) return;
  delete static_cast<HostOwner*>(o->_owner);
  memset(o, 0, sizeof(*o));
}

// ---- host-buffer entry points: upload, run, download --------------------------------
struct SegList {
  std::vector<ii2_seg*> v;
  ~SegList() {
    for (ii2_seg* g : v) delete g;
  }
};

int ii2_merge(const ii2_seg_view* segs, int nseg, const uint32_t* removed_sorted, uint64_t nrem,
              uint32_t flags, ii2_merge_out* out) {
  if (!out || nseg < 0 || (nseg && !segs)) return II2_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  II2_TRY(ctx_require());
  SegList list;
  for (int i = 0; i < nseg; i++) {
    ii2_seg* g = nullptr;
    II2_TRY(ii2_seg_upload(&segs[i], &g));
    list.v.push_back(g);
  }
  ii2_removed* rem = nullptr;
  if (nrem) II2_TRY(ii2_removed_upload(removed_sorted, nrem, &rem));
  std::unique_ptr<ii2_removed> rem_guard(rem);
  ii2_result* res = nullptr;
  II2_TRY(ii2_merge_dev(list.v.data(), nseg, rem,
                        II2_RESULT_ENCODED | ((flags & II2_MERGE_WANT_DECODED) ? II2_RESULT_DECODED : 0u),
                        &res));
  std::unique_ptr<ii2_result> res_guard(res);
  return ii2_result_download_merge(res, flags, out);
}

int ii2_read_range(const ii2_seg_view* segs, int nseg, const uint8_t* min, size_t minlen,
                   const uint8_t* max, size_t maxlen, const uint32_t* removed_sorted,
                   uint64_t nrem, ii2_read_out* out) {
  if (!out || nseg < 0 || (nseg && !segs)) return II2_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  II2_TRY(ctx_require());
  SegList list;
  for (int i = 0; i < nseg; i++) {
    ii2_seg* g = nullptr;
    II2_TRY(ii2_seg_upload(&segs[i], &g));
    list.v.push_back(g);
  }
  ii2_removed* rem = nullptr;
  if (removed_sorted) II2_TRY(ii2_removed_upload(removed_sorted, nrem, &rem));
  std::unique_ptr<ii2_removed> rem_guard(rem);
  ii2_result* res = nullptr;
  II2_TRY(ii2_read_range_dev(list.v.data(), nseg, min, minlen, max, maxlen, rem, &res));
  std::unique_ptr<ii2_result> res_guard(res);
  return ii2_result_download_read(res, out);
}

}  // extern "C"
