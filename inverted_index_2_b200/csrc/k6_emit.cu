// k6_emit.cu — emit the merged segment / read result from the union results.
//
// Replaces the writer half of the merge loop: empty-term drop (shard.go:192-194),
// Writer.Append (file/writer.go:32-59: FST output = running valuesOffset, `_val` = one
// intcomp stream per term, little-endian words, no framing) and, for reads, the TermValues
// the iterator would yield.  One CTA per K1 bucket; positions are walked in chunks of 256,
// a block scan over the surviving heads gives every term its output slots (bucket bases come
// from the bucket-level scan in k2_union), then warps copy term bytes, postings and encode.
#include "intcomp.cuh"
#include "union.cuh"

namespace ii2 {

constexpr int K6_THREADS = 256;
constexpr int K6_WARPS = K6_THREADS / 32;

struct K6Args {
  const SegDesc* segs;
  int k;
  const uint32_t* bk_pos;
  const uint32_t* ord_inst;
  const uint16_t* gsz;
  const uint32_t* tmp_post;
  const uint32_t* g_cnt;
  const uint32_t* g_enc;
  const uint64_t* g_off;
  const uint64_t* bk_out;  // [4][nb1] exclusive prefixes
  uint32_t nb1;
  int want_decoded, want_enc, keep_empty;
  uint8_t* o_term_bytes;
  uint32_t* o_term_off;
  uint32_t* o_post;
  uint64_t* o_post_off;
  uint32_t* o_val_words;
  uint64_t* o_val_off;
};

struct K6Work {
  const uint8_t* tsrc;
  uint32_t tlen;
  uint32_t tdst;
  uint32_t cnt;
  uint64_t src_off;
  uint64_t post_dst;
  uint64_t enc_dst;
};

__global__ void __launch_bounds__(K6_THREADS) k6_emit_kernel(const K6Args a) {
  __shared__ K6Work s_work[K6_THREADS];
  __shared__ uint64_t s_ws64[K6_WARPS + 2];
  __shared__ uint32_t s_ws32[K6_WARPS + 2];
  __shared__ uint32_t s_stage[K6_WARPS][intcomp::kStageWords];
  const uint32_t tid = threadIdx.x, b = blockIdx.x;
  const uint32_t p0 = a.bk_pos[b], p1 = a.bk_pos[b + 1];
  uint64_t run_t = a.bk_out[0ull * a.nb1 + b];
  uint64_t run_tb = a.bk_out[1ull * a.nb1 + b];
  uint64_t run_p = a.bk_out[2ull * a.nb1 + b];
  uint64_t run_e = a.bk_out[3ull * a.nb1 + b];
  for (uint32_t q = p0; q < p1; q += K6_THREADS) {
    const uint32_t p = q + tid;
    uint32_t cnt = 0, enc = 0, tl = 0;
    const uint8_t* tsrc = nullptr;
    uint32_t surv = 0;
    if (p < p1 && a.gsz[p] != 0) {
      cnt = a.g_cnt[p];
      if (cnt || a.keep_empty) {
        surv = 1;
        enc = a.g_enc[p];
        int s;
        uint32_t idx;
        locate_instance(a.segs, a.k, a.ord_inst[p], s, idx);
        const uint32_t o = a.segs[s].toff[idx];
        tl = a.segs[s].toff[idx + 1] - o;
        tsrc = a.segs[s].tb + o;
      }
    }
    uint32_t n_surv;
    const uint32_t ex_t = block_exclusive_scan(surv, s_ws32, n_surv);
    uint32_t tot_tb;
    const uint32_t ex_tb = block_exclusive_scan(tl, s_ws32, tot_tb);
    uint64_t tot_p, tot_e;
    const uint64_t ex_p = block_exclusive_scan((uint64_t)cnt, s_ws64, tot_p);
    const uint64_t ex_e = block_exclusive_scan((uint64_t)enc, s_ws64, tot_e);
    if (surv) {
      const uint64_t t = run_t + ex_t;
      K6Work w;
      w.tsrc = tsrc;
      w.tlen = tl;
      w.tdst = (uint32_t)(run_tb + ex_tb);
      w.cnt = cnt;
      w.src_off = a.g_off[p];
      w.post_dst = run_p + ex_p;
      w.enc_dst = run_e + ex_e;
      s_work[ex_t] = w;
      a.o_term_off[t] = w.tdst;
      if (a.want_decoded) a.o_post_off[t] = w.post_dst;
      if (a.want_enc) a.o_val_off[t] = 4ull * w.enc_dst;
    }
    __syncthreads();
    for (uint32_t h = warp_id(); h < n_surv; h += K6_WARPS) {
      const K6Work w = s_work[h];
      const unsigned lane = lane_id();
      for (uint32_t i = lane; i < w.tlen; i += 32) a.o_term_bytes[w.tdst + i] = w.tsrc[i];
      const uint32_t* src = a.tmp_post + w.src_off;
      if (a.want_decoded)
        for (uint32_t i = lane; i < w.cnt; i += 32) a.o_post[w.post_dst + i] = src[i];
      if (a.want_enc) intcomp::enc_emit_warp(src, w.cnt, a.o_val_words + w.enc_dst, s_stage[warp_id()]);
    }
    run_t += n_surv;
    run_tb += tot_tb;
    run_p += tot_p;
    run_e += tot_e;
    __syncthreads();
  }
  // terminal offsets, written once by the last bucket
  if (b == gridDim.x - 1 && tid == 0) {
    a.o_term_off[run_t] = (uint32_t)run_tb;
    if (a.want_decoded) a.o_post_off[run_t] = run_p;
  }
}

int k6_emit(const MergePlan& plan, const UnionOut& u, bool want_decoded, bool want_enc,
            EmitOut& out, cudaStream_t s) {
  const uint64_t T = u.h_totals[0], TB = u.h_totals[1], P = u.h_totals[2], E = u.h_totals[3];
  if (TB >= (1ull << 32)) {
    set_last_error("merged term dictionary exceeds 4 GiB of term bytes");
    return II2_ERR_UNSUPPORTED;
  }
  II2_TRY(out.term_bytes.alloc(TB, s, 32));
  II2_TRY(out.term_off.alloc(T + 1, s));
  if (want_decoded) {
    II2_TRY(out.post.alloc(P, s));
    II2_TRY(out.post_off.alloc(T + 1, s));
  }
  if (want_enc) {
    II2_TRY(out.val_words.alloc(E, s));
    II2_TRY(out.val_off.alloc(T, s));
  }
  K6Args a;
  a.segs = plan.segs;
  a.k = plan.k;
  a.bk_pos = plan.bk_pos.p;
  a.ord_inst = plan.ord_inst.p;
  a.gsz = plan.gsz.p;
  a.tmp_post = u.tmp_post.p;
  a.g_cnt = u.g_cnt.p;
  a.g_enc = u.g_enc.p;
  a.g_off = u.g_off.p;
  a.bk_out = u.bk_out.p;
  a.nb1 = plan.n_buckets + 1;
  a.want_decoded = want_decoded;
  a.want_enc = want_enc;
  a.keep_empty = u.keep_empty ? 1 : 0;
  a.o_term_bytes = out.term_bytes.p;
  a.o_term_off = out.term_off.p;
  a.o_post = out.post.p;
  a.o_post_off = out.post_off.p;
  a.o_val_words = out.val_words.p;
  a.o_val_off = out.val_off.p;
  ProfScope scope("k6_emit", s);
  k6_emit_kernel<<<plan.n_buckets, K6_THREADS, 0, s>>>(a);
  II2_LAUNCHED();
  return II2_OK;
}

}  // namespace ii2
