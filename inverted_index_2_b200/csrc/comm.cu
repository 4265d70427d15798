// comm.cu — the cross-shard exchange behind the C-ABI (include/ii2.h, "cross-shard exchange").
//
// Replaces, when the shards of one index live on several GPUs (one process per GPU, contiguous
// shard-key ranges per rank, so rank order == term order):
//   - the ordered concatenation of shard streams of InvertedIndex.Read
//     (inverted_index.go:330-338): ii2_read_gather — ONE all-gather of a 32-byte size record,
//     ONE group of ncclSend / ncclRecv that moves the four flat arrays of every rank straight
//     from its result to their final place on the root, and one kernel that rebases the offsets;
//   - the union of the per-shard maps of PrefixSearch under its mutex + the final slices.Sort /
//     slices.Compact (inverted_index.go:274-292): ii2_prefix_gather — the per-prefix value lists
//     of every rank land back to back on the root and the heavy-term union kernels
//     (k12_union.cu, sort forced) make every prefix sorted-unique again.
// Compaction itself needs no exchange: shards never interact (shard.go:19-20).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library must load on hosts without
// it, and inside a process that already carries torch's NCCL the same instance is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <memory>
#include <mutex>
#include <vector>

#include "handles.cuh"

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    bool all = true;
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(h, name);
      if (!p) all = false;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = all;
  });
  return api;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
Comm g_comm;
std::mutex g_comm_mu;

#define II2_NCCL_TRY(expr)                                                              \
  do {                                                                                  \
    ncclResult_t _r = (expr);                                                           \
    if (_r != ncclSuccess) {                                                            \
      set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(_r)); \
      return II2_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

constexpr int kMaxWorld = 64;

struct Bases {
  uint64_t T[kMaxWorld + 1], TB[kMaxWorld + 1], P[kMaxWorld + 1];
  int world;
};

// term_off / post_off of rank r's part, received as they were (relative to the part), shifted
// to the concatenation; the terminal entries close the arrays
__global__ void __launch_bounds__(256)
k_rebase_parts(uint32_t* __restrict__ toff, uint64_t* __restrict__ poff, const Bases b) {
  const uint64_t T = b.T[b.world];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= T;
       i += (uint64_t)gridDim.x * blockDim.x) {
    if (i == T) {
      toff[T] = (uint32_t)b.TB[b.world];
      poff[T] = b.P[b.world];
      continue;
    }
    int r = 0;
    while (b.T[r + 1] <= i) r++;
    toff[i] += (uint32_t)b.TB[r];
    poff[i] += b.P[r];
  }
}

// one (prefix, rank) source of the union on the root: where rank r's values of prefix p landed
__global__ void __launch_bounds__(256)
k_prefix_sources(const uint64_t* __restrict__ voff_all /* [world][np+1] */, const uint32_t* __restrict__ vals,
                 const uint64_t* __restrict__ vbase /* [world] */, uint32_t np, int world,
                 uint64_t* __restrict__ src_ptr, uint32_t* __restrict__ src_len, GroupIn* __restrict__ gin,
                 uint32_t* __restrict__ rec, uint32_t* __restrict__ flags) {
  const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (uint64_t)np * world) return;
  const uint32_t p = (uint32_t)(id / world);
  const int r = (int)(id % world);
  const uint64_t* vo = voff_all + (uint64_t)r * (np + 1);
  uint64_t len = vo[p + 1] - vo[p];
  if (len > 0xFFFFFFFFull) {
    atomicExch(&flags[0], 1u);
    len = 0;
  }
  src_ptr[id] = reinterpret_cast<uint64_t>(vals + vbase[r] + vo[p]);
  src_len[id] = (uint32_t)len;
  if (r == 0) {
    GroupIn g;
    g.inst = 0;
    g.tlen = 0;
    g.src = (uint32_t)id;
    g.c = (uint32_t)world;
    g.L = 0xFFFFFFFFu;
    g.pst = 0;
    g.eslot = 0;
    g.pad = 0;
    gin[p] = g;
    rec[p] = p;
  }
}

__global__ void __launch_bounds__(256)
k_prefix_counts(const GroupRec* __restrict__ recs, uint32_t np, uint64_t* __restrict__ cnt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < np) cnt[p] = recs[p].cnt;
  if (p == np) cnt[np] = 0;
}

__global__ void __launch_bounds__(256)
k_prefix_place(const GroupRec* __restrict__ recs, const uint64_t* __restrict__ off,
               uint32_t* __restrict__ out) {
  const GroupRec r = recs[blockIdx.y];
  const uint32_t* src = reinterpret_cast<const uint32_t*>(r.dec);
  uint32_t* dst = out + off[blockIdx.y];
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < r.cnt; i += (uint64_t)gridDim.x * 256)
    dst[i] = src[i];
}

__global__ void __launch_bounds__(256)
k_or_matched(const uint8_t* __restrict__ all /* [world][np] */, uint32_t np, int world,
             uint8_t* __restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  uint8_t m = 0;
  for (int r = 0; r < world; r++) m |= all[(uint64_t)r * np + p];
  out[p] = m ? 1 : 0;
}

int comm_ready() {
  if (!g_comm.comm) {
    set_last_error("ii2_comm_init has not been called");
    return II2_ERR_INVALID;
  }
  return II2_OK;
}

int read_gather_impl(const ii2_result* local, int root, ii2_result** out, cudaStream_t s) {
  NcclApi& n = nccl();
  const int world = g_comm.world, rank = g_comm.rank;
  const bool recv = root < 0 || rank == root;
  // ---- sizes of every rank: one all-gather of {terms, term bytes, postings, -}
  DevBuf<uint64_t> d_sz;
  II2_TRY(d_sz.alloc_scratch(4 * (size_t)(world + 1), s));
  uint64_t* h = static_cast<uint64_t*>(pinned_alloc(8 * 4 * (size_t)(world + 1)));
  if (!h) return II2_ERR_NOMEM;
  struct Guard {
    void* p;
    ~Guard() { pinned_free(p); }
  } guard{h};
  uint64_t* mine = h + 4 * (size_t)world;
  mine[0] = local->T;
  mine[1] = local->TB;
  mine[2] = local->P;
  mine[3] = 0;
  II2_TRY(small_copy(d_sz.p + 4 * (size_t)world, mine, 32, s));
  II2_NCCL_TRY(n.AllGather(d_sz.p + 4 * (size_t)world, d_sz.p, 4, ncclUint64, g_comm.comm, s));
  II2_TRY(small_copy(h, d_sz.p, 32 * (size_t)world, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  Bases b;
  b.world = world;
  b.T[0] = b.TB[0] = b.P[0] = 0;
  for (int r = 0; r < world; r++) {
    b.T[r + 1] = b.T[r] + h[4 * r];
    b.TB[r + 1] = b.TB[r] + h[4 * r + 1];
    b.P[r + 1] = b.P[r] + h[4 * r + 2];
  }
  if (b.TB[world] >= (1ull << 32) || b.T[world] >= 0xFFFFFFFFull) {
    set_last_error("gathered read exceeds 4 GiB of term bytes / 2^32-2 terms");
    return II2_ERR_UNSUPPORTED;
  }
  std::unique_ptr<ii2_result> res(new ii2_result());
  res->has_dec = true;
  EmitOut& o = res->out;
  if (recv) {
    res->T = b.T[world];
    res->TB = b.TB[world];
    res->P = b.P[world];
    II2_TRY(o.term_bytes.alloc(res->TB, s, 32));
    II2_TRY(o.term_off.alloc(res->T + 1, s, 16));
    II2_TRY(o.post.alloc(res->P, s, 16));
    II2_TRY(o.post_off.alloc(res->T + 1, s, 16));
  } else {
    II2_TRY(o.term_bytes.alloc(0, s, 32));
    II2_TRY(o.term_off.alloc(1, s));
    II2_TRY(o.post.alloc(0, s));
    II2_TRY(o.post_off.alloc(1, s));
    II2_CUDA_TRY(cudaMemsetAsync(o.term_off.p, 0, 4, s));
    II2_CUDA_TRY(cudaMemsetAsync(o.post_off.p, 0, 8, s));
  }
  // ---- one exchange: every array of every rank straight to its final place
  const EmitOut& l = local->out;
  auto put = [&](int r, const void* src, void* dst, uint64_t bytes) -> int {  // rank r's piece
    if (!bytes) return II2_OK;
    if (r == rank) {
      if (recv) II2_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
      return II2_OK;
    }
    II2_NCCL_TRY(n.Recv(dst, bytes, ncclUint8, r, g_comm.comm, s));
    return II2_OK;
  };
  II2_NCCL_TRY(n.GroupStart());  // (closed below whatever happens inside: an open group poisons the communicator)
  int rc = II2_OK;
  for (int dst_rank = 0; dst_rank < world && rc == II2_OK; dst_rank++) {
    if (!(root < 0 || dst_rank == root) || dst_rank == rank) continue;
    // my pieces to the receiving rank (the same order on both sides)
    const uint64_t T = local->T;
    auto send = [&](const void* p, uint64_t bytes) -> int {
      if (!bytes) return II2_OK;
      II2_NCCL_TRY(n.Send(p, bytes, ncclUint8, dst_rank, g_comm.comm, s));
      return II2_OK;
    };
    if (rc == II2_OK) rc = send(l.term_bytes.p, local->TB);
    if (rc == II2_OK) rc = send(l.term_off.p, T * 4);
    if (rc == II2_OK) rc = send(l.post.p, local->P * 4);
    if (rc == II2_OK) rc = send(l.post_off.p, T * 8);
  }
  if (recv) {
    for (int r = 0; r < world && rc == II2_OK; r++) {
      const uint64_t T = h[4 * r], TBr = h[4 * r + 1], Pr = h[4 * r + 2];
      const bool me = r == rank;
      if (rc == II2_OK) rc = put(r, me ? l.term_bytes.p : nullptr, o.term_bytes.p + b.TB[r], TBr);
      if (rc == II2_OK) rc = put(r, me ? l.term_off.p : nullptr, o.term_off.p + b.T[r], T * 4);
      if (rc == II2_OK) rc = put(r, me ? l.post.p : nullptr, o.post.p + b.P[r], Pr * 4);
      if (rc == II2_OK) rc = put(r, me ? l.post_off.p : nullptr, o.post_off.p + b.T[r], T * 8);
    }
  }
  {
    const ncclResult_t ge = n.GroupEnd();
    II2_TRY(rc);
    II2_NCCL_TRY(ge);
  }
  if (recv) {
    const unsigned grid = (unsigned)std::min<uint64_t>(div_up(res->T + 1, 256), 1184);
    k_rebase_parts<<<grid, 256, 0, s>>>(o.term_off.p, o.post_off.p, b);
    II2_LAUNCHED();
  }
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  *out = res.release();
  return II2_OK;
}

int prefix_gather_impl(const ii2_prefix_out* local, int root, ii2_prefix_out* out, cudaStream_t s) {
  NcclApi& n = nccl();
  const int world = g_comm.world, rank = g_comm.rank;
  const bool recv = root < 0 || rank == root;
  const uint32_t np = (uint32_t)local->n_prefixes;
  const uint64_t nv = np ? local->value_off[np] : 0;
  // ---- every rank's value offsets and matched flags: fixed size, two all-gathers
  DevBuf<uint64_t> d_voff;
  DevBuf<uint8_t> d_m;
  const size_t row = (size_t)np + 1;
  II2_TRY(d_voff.alloc_scratch(row * (size_t)(world + 1), s));
  II2_TRY(d_m.alloc_scratch(((size_t)np + 16) * (size_t)(world + 1), s));
  uint64_t* h_voff = static_cast<uint64_t*>(pinned_alloc(8 * row * (size_t)world + 64));
  if (!h_voff) return II2_ERR_NOMEM;
  struct Guard {
    void* p;
    ~Guard() { pinned_free(p); }
  } guard{h_voff};
  uint64_t* my_voff = d_voff.p + row * (size_t)world;
  uint8_t* my_m = d_m.p + (size_t)np * world;
  if (np) {
    II2_CUDA_TRY(cudaMemcpyAsync(my_voff, local->value_off, row * 8, cudaMemcpyHostToDevice, s));
    II2_CUDA_TRY(cudaMemcpyAsync(my_m, local->matched, np, cudaMemcpyHostToDevice, s));
  } else {
    II2_CUDA_TRY(cudaMemsetAsync(my_voff, 0, 8, s));
  }
  II2_NCCL_TRY(n.AllGather(my_voff, d_voff.p, row, ncclUint64, g_comm.comm, s));
  if (np) II2_NCCL_TRY(n.AllGather(my_m, d_m.p, np, ncclUint8, g_comm.comm, s));
  II2_CUDA_TRY(cudaMemcpyAsync(h_voff, d_voff.p, 8 * row * (size_t)world, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  std::vector<uint64_t> vbase(world + 1, 0);
  for (int r = 0; r < world; r++) vbase[r + 1] = vbase[r] + h_voff[row * r + np];
  // ---- the values: one group of sends / receives, rank r's block at vbase[r]
  DevBuf<uint32_t> d_vals, d_mine;
  II2_TRY(d_vals.alloc_scratch(recv ? vbase[world] : 0, s));
  II2_TRY(d_mine.alloc_scratch(nv, s));
  if (nv) II2_CUDA_TRY(cudaMemcpyAsync(d_mine.p, local->values, nv * 4, cudaMemcpyHostToDevice, s));
  II2_NCCL_TRY(n.GroupStart());
  {
    int rc = II2_OK;
    auto xfer = [&](bool send, void* p, uint64_t bytes, int peer) -> int {
      if (send)
        II2_NCCL_TRY(n.Send(p, bytes, ncclUint8, peer, g_comm.comm, s));
      else
        II2_NCCL_TRY(n.Recv(p, bytes, ncclUint8, peer, g_comm.comm, s));
      return II2_OK;
    };
    for (int dst_rank = 0; dst_rank < world && rc == II2_OK; dst_rank++) {
      if (!(root < 0 || dst_rank == root) || dst_rank == rank || !nv) continue;
      rc = xfer(true, d_mine.p, nv * 4, dst_rank);
    }
    if (recv) {
      for (int r = 0; r < world && rc == II2_OK; r++) {
        const uint64_t cnt = vbase[r + 1] - vbase[r];
        if (!cnt) continue;
        if (r == rank) {
          if (cudaMemcpyAsync(d_vals.p + vbase[r], d_mine.p, cnt * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
            set_last_error("prefix gather: local copy failed");
            rc = II2_ERR_CUDA;
          }
        } else {
          rc = xfer(false, d_vals.p + vbase[r], cnt * 4, r);
        }
      }
    }
    const ncclResult_t ge = n.GroupEnd();  // always closed
    II2_TRY(rc);
    II2_NCCL_TRY(ge);
  }
  std::unique_ptr<HostOwner> own(new HostOwner());
  out->n_prefixes = np;
  out->matched = own->alloc<uint8_t>((size_t)np + 1);
  out->value_off = own->alloc<uint64_t>(row);
  if (!out->matched || !out->value_off) return II2_ERR_NOMEM;
  if (!recv || np == 0) {
    memset(out->matched, 0, (size_t)np + 1);
    memset(out->value_off, 0, row * 8);
    out->values = own->alloc<uint32_t>(0);
    II2_CUDA_TRY(cudaStreamSynchronize(s));
    out->_owner = own.release();
    return II2_OK;
  }
  // ---- per prefix: the sorted-unique union of its `world` lists (inverted_index.go:289-292)
  const uint64_t pairs = (uint64_t)np * world;
  DevBuf<uint64_t> src_ptr, d_vbase, d_off, d_tot;
  DevBuf<uint32_t> src_len, rec, flags, large_tmp, large_enc;
  DevBuf<GroupIn> gin;
  DevBuf<GroupRec> recs;
  DevBuf<uint8_t> d_mout;
  II2_TRY(src_ptr.alloc_scratch(pairs, s));
  II2_TRY(src_len.alloc_scratch(pairs, s));
  II2_TRY(rec.alloc_scratch(np, s));
  II2_TRY(flags.alloc_scratch(2, s));
  II2_TRY(gin.alloc_scratch(np, s));
  II2_TRY(recs.alloc_scratch(np, s));
  II2_TRY(d_vbase.alloc_scratch(world + 1, s));
  II2_TRY(d_off.alloc_scratch(row, s));
  II2_TRY(d_tot.alloc_scratch(1, s));
  II2_TRY(d_mout.alloc_scratch(np, s));
  II2_CUDA_TRY(cudaMemsetAsync(flags.p, 0, 8, s));
  uint64_t* h_small = static_cast<uint64_t*>(pinned_alloc(8 * (size_t)(world + 4)));
  if (!h_small) return II2_ERR_NOMEM;
  Guard g2{h_small};
  for (int r = 0; r <= world; r++) h_small[r] = vbase[r];
  II2_TRY(small_copy(d_vbase.p, h_small, 8 * (size_t)(world + 1), s));
  k_prefix_sources<<<div_up(pairs, 256), 256, 0, s>>>(d_voff.p, d_vals.p, d_vbase.p, np, world, src_ptr.p,
                                                      src_len.p, gin.p, rec.p, flags.p);
  II2_LAUNCHED();
  LargeArgs la;
  la.rec = rec.p;
  la.bucket = nullptr;
  la.gin = gin.p;
  la.src_ptr = src_ptr.p;
  la.src_len = src_len.p;
  la.recs = recs.p;
  la.rem.sorted = nullptr;
  la.rem.n = 0;
  la.rem.bitmap = nullptr;
  la.rem.bitmap_bits = 0;
  la.want_enc = 0;
  la.keep_empty = 1;
  la.always_sort = 1;
  la.presorted = nullptr;
  la.bk_raw = nullptr;
  la.nb1 = 0;
  II2_TRY(k2_large_run(la, np, large_tmp, large_enc, s));
  k_prefix_counts<<<div_up((uint64_t)np + 1, 256), 256, 0, s>>>(recs.p, np, d_off.p);
  II2_LAUNCHED();
  II2_TRY(exclusive_scan_u64(d_off.p, row, d_tot.p, s));
  k_or_matched<<<div_up(np, 256), 256, 0, s>>>(d_m.p, np, world, d_mout.p);
  II2_LAUNCHED();
  II2_TRY(small_copy(h_small, d_tot.p, 8, s));
  II2_TRY(small_copy(h_small + 1, flags.p, 8, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  if ((uint32_t)h_small[1]) {
    set_last_error("prefix gather: one rank holds more than 2^32-1 values under a prefix");
    return II2_ERR_UNSUPPORTED;
  }
  const uint64_t total = h_small[0];
  DevBuf<uint32_t> d_out;
  II2_TRY(d_out.alloc_scratch(total, s));
  if (total) {
    const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(1024, (total / np + 2047) / 2048));
    for (uint32_t y0 = 0; y0 < np; y0 += 32768) {
      const uint32_t ny = std::min<uint32_t>(32768, np - y0);
      k_prefix_place<<<dim3(gx, ny), 256, 0, s>>>(recs.p + y0, d_off.p + y0, d_out.p);
      II2_LAUNCHED();
    }
  }
  out->values = own->alloc<uint32_t>(total);
  if (!out->values) return II2_ERR_NOMEM;
  if (total) II2_CUDA_TRY(cudaMemcpyAsync(out->values, d_out.p, total * 4, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaMemcpyAsync(out->value_off, d_off.p, row * 8, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaMemcpyAsync(out->matched, d_mout.p, np, cudaMemcpyDeviceToHost, s));
  II2_CUDA_TRY(cudaStreamSynchronize(s));
  out->_owner = own.release();
  return II2_OK;
}

}  // namespace

extern "C" {

int ii2_comm_unique_id(uint8_t* id) {
  if (!id) return II2_ERR_INVALID;
  if (!nccl().ok) {
    set_last_error("libnccl.so.2 could not be loaded");
    return II2_ERR_UNSUPPORTED;
  }
  static_assert(sizeof(ncclUniqueId) <= II2_COMM_ID_BYTES, "unique id fits the C-ABI buffer");
  ncclUniqueId u;
  II2_NCCL_TRY(nccl().GetUniqueId(&u));
  memset(id, 0, II2_COMM_ID_BYTES);
  memcpy(id, &u, sizeof(u));
  return II2_OK;
}

int ii2_comm_init(const uint8_t* id, int rank, int world) {
  if (!id || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return II2_ERR_INVALID;
  II2_TRY(ctx_require());
  if (!nccl().ok) {
    set_last_error("libnccl.so.2 could not be loaded");
    return II2_ERR_UNSUPPORTED;
  }
  std::lock_guard<std::mutex> lock(g_comm_mu);
  if (g_comm.comm) {
    set_last_error("ii2_comm_init called twice (ii2_comm_shutdown first)");
    return II2_ERR_INVALID;
  }
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t c = nullptr;
  II2_NCCL_TRY(nccl().CommInitRank(&c, world, u, rank));
  g_comm.comm = c;
  g_comm.rank = rank;
  g_comm.world = world;
  return II2_OK;
}

int ii2_comm_info(int* rank, int* world) {
  if (rank) *rank = g_comm.comm ? g_comm.rank : 0;
  if (world) *world = g_comm.comm ? g_comm.world : 0;
  return II2_OK;
}

void ii2_comm_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_comm_mu);
  if (g_comm.comm) {
    nccl().CommDestroy(g_comm.comm);
    g_comm = Comm();
  }
}

int ii2_read_gather(const ii2_result* local, int root, ii2_result** gathered) {
  if (!local || !gathered) return II2_ERR_INVALID;
  *gathered = nullptr;
  II2_TRY(ctx_require());
  II2_TRY(comm_ready());
  if (root >= g_comm.world || !local->has_dec) return II2_ERR_INVALID;
  cudaStream_t s = cur_stream();
  const int rc = read_gather_impl(local, root, gathered, s);
  if (rc != II2_OK) cudaStreamSynchronize(s);
  arena_reset(s);
  return rc;
}

int ii2_prefix_gather(const ii2_prefix_out* local, int root, ii2_prefix_out* merged) {
  if (!local || !merged) return II2_ERR_INVALID;
  memset(merged, 0, sizeof(*merged));
  II2_TRY(ctx_require());
  II2_TRY(comm_ready());
  if (root >= g_comm.world || (local->n_prefixes && (!local->value_off || !local->matched)))
    return II2_ERR_INVALID;
  cudaStream_t s = cur_stream();
  const int rc = prefix_gather_impl(local, root, merged, s);
  if (rc != II2_OK) {
    cudaStreamSynchronize(s);
    memset(merged, 0, sizeof(*merged));
  }
  arena_reset(s);
  return rc;
}

}  // extern "C"
