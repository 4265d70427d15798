// codec.cuh — host-callable entry points of the device codec (k3a_intcomp.cu).
#pragma once
#include "runtime.cuh"

namespace ii2 {
// words/offsets are device pointers; out/out_off are allocated here.  Synchronises `s`.
// scratch_out: the outputs are intermediates of the calling entry point (arena) instead of
// buffers that outlive it.  Temporaries always come from the arena: the caller resets it.
int intcomp_decode_dev(const uint32_t* d_words, const uint64_t* d_woff, uint64_t nlists,
                       DevBuf<uint32_t>& out, DevBuf<uint64_t>& out_off, uint64_t* total_out,
                       cudaStream_t s, bool scratch_out);
// nvals_hint = off[nlists] - off[0] (known to the caller; sizes the long-list scratch)
int intcomp_encode_dev(const uint32_t* d_in, const uint64_t* d_off, uint64_t nlists,
                       uint64_t nvals_hint, DevBuf<uint32_t>& words, DevBuf<uint64_t>& woff,
                       uint64_t* total_words, cudaStream_t s);
int val_offsets_to_word_offsets(const uint64_t* d_val_off, uint64_t n, uint64_t val_size,
                                DevBuf<uint64_t>& woff, cudaStream_t s);
}  // namespace ii2
