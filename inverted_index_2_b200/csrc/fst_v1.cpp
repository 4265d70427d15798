// fst_v1.cpp — reader and writer for the `<key>_fst` term dictionaries: blevesearch/vellum
// v1.0.10 FST files, encoding version 1 (go.mod:7; written through vellum.New at
// file/writer.go:104-129 and Insert at :35,43; read through vellum.Open / Iterator at
// file/reader.go:139-151 and :48-50).  SURVEY.md §8(f) row 1: with this, the library consumes
// and produces real segment directories instead of pre-flattened views.
//
// Host-side C++ (the FST is a sequential pointer structure in the reference too — not a GPU
// kernel); no CUDA call is made here, so these entry points also work without a device.
//
// PARITY UNPINNED: the vellum module is not in /root/reference and cannot be fetched, and the
// reference's tests hold no `_fst` bytes.  The format below is a restatement of vellum's
// published v1 encoding (encoder_v1.go / decoder_v1.go / builder.go / registry.go /
// common.go); reader and writer are checked against each other and against hand-derived byte
// vectors (tests/test_fst.py), not against bytes produced by Go.
//
// File layout (little-endian):
//   header   u64 version = 1, u64 type = 0
//   states   written children-first; a state's ADDRESS is the offset of its LAST byte and it
//            is decoded backwards from there; transition targets are stored as the distance
//            from the state's FIRST byte (0 = the implicit "final, no transitions, no output"
//            state, which is never written)
//            - one transition, not final:  [out?][delta][pack] [in?] hdr
//                hdr = 0x80 | 0x40 (target is the state written just before, no out/delta/pack)
//                      | common-input code (1..63; 0 = explicit input byte below the header)
//            - anything else:  [final out?][outs, reversed][deltas, reversed][ins, reversed]
//                              pack [count?] hdr
//                hdr = 0x40 (final) | count (1..63; 0 = count byte below; 256 is stored as 1)
//              pack = delta bytes << 4 | output bytes (0 = no outputs at all)
//   footer   u64 number of keys, u64 root address
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/ii2.h"

namespace {

// ---------------------------------------------------------------- common inputs (common.go)
// Inputs ranked by frequency in URLs-and-words corpora (the table vellum took from
// BurntSushi/fst); only ranks 0..62 get a one-byte code (rank + 1) inside the state header.
const uint8_t kCommonInv[63] = {
    't', 'e', '/', 'o', 'a', 's', 'r', 'i', 'p', 'c', 'n', 'w', '.', 'h', 'l', 'm',
    '-', 'd', 'u', '0', '1', '2', 'g', '=', ':', 'b', 'f', '3', 'y', '5', '&', '_',
    '4', 'v', '9', '6', '7', '8', 'k', '%', '?', 'x', 'C', 'D', 'A', 'S', 'F', 'I',
    'B', 'E', 'j', 'P', 'T', 'z', 'R', 'N', 'M', '+', 'L', 'O', 'q', 'H', 'G'};

struct CommonTable {
  uint8_t code[256];
  CommonTable() {
    memset(code, 0, sizeof(code));
    for (int i = 0; i < 63; i++) code[kCommonInv[i]] = (uint8_t)(i + 1);
  }
};
const CommonTable kCommon;

constexpr uint8_t kOneTransition = 0x80, kTransitionNext = 0x40, kStateFinal = 0x40;
constexpr uint8_t kMaxCommon = 63, kMaxNumTrans = 63;
constexpr uint64_t kEmptyAddr = 0, kNoneAddr = 1;
constexpr size_t kHeader = 16, kFooter = 16;

int packed_size(uint64_t n) {
  int s = 1;
  while (s < 8 && (n >> (8 * s)) != 0) s++;
  return s;
}

uint64_t read_packed(const uint8_t* p, int n) {
  uint64_t v = 0;
  for (int i = n - 1; i >= 0; i--) v = (v << 8) | p[i];
  return v;
}

uint64_t read_u64(const uint8_t* p) { return read_packed(p, 8); }

// ---------------------------------------------------------------- writer (builder.go)
struct Trans {
  uint8_t in;
  uint64_t out;
  uint64_t addr;
  bool operator==(const Trans& o) const { return in == o.in && out == o.out && addr == o.addr; }
};

struct Node {
  bool final = false;
  uint64_t final_output = 0;
  std::vector<Trans> trans;
  bool equiv(const Node& o) const {
    return final == o.final && final_output == o.final_output && trans == o.trans;
  }
};

struct Unfinished {
  Node node;
  bool has_last = false;
  uint8_t last_in = 0;
  uint64_t last_out = 0;
  void last_compiled(uint64_t addr) {
    if (has_last) {
      node.trans.push_back(Trans{last_in, last_out, addr});
      has_last = false;
      last_out = 0;
    }
  }
  void add_output_prefix(uint64_t prefix) {
    if (node.final) node.final_output += prefix;
    for (Trans& t : node.trans) t.out += prefix;
    if (has_last) last_out += prefix;
  }
};

// registry.go: suffix sharing through a bounded table of recently compiled states —
// `table_size` buckets of `mru` cells, most recently used first; a miss evicts the bucket's
// least recently used cell.  vellum's defaults (builder.go: 10 000 x 2) are kept so that the
// same states get shared as in a file written by Go.
struct Registry {
  struct Cell {
    bool used = false;
    uint64_t addr = 0;
    Node node;
  };
  size_t table_size, mru;
  std::vector<Cell> cells;
  Registry(size_t ts, size_t m) : table_size(ts), mru(m), cells(ts * m) {}
  size_t hash(const Node& n) const {
    const uint64_t prime = 1099511628211ull;
    uint64_t h = 14695981039346656037ull;
    h = (h ^ (n.final ? 1u : 0u)) * prime;
    h = (h ^ n.final_output) * prime;
    for (const Trans& t : n.trans) {
      h = (h ^ (uint64_t)t.in) * prime;
      h = (h ^ t.out) * prime;
      h = (h ^ t.addr) * prime;
    }
    return (size_t)(h % table_size);
  }
  // found -> true + addr; else the node is installed as the bucket's most recent cell and a
  // pointer to its address slot is returned for the caller to fill
  bool entry(const Node& n, uint64_t& addr, uint64_t** slot) {
    Cell* r = &cells[hash(n) * mru];
    for (size_t i = 0; i < mru; i++) {
      if (r[i].used && r[i].node.equiv(n)) {
        addr = r[i].addr;
        for (size_t j = i; j > 0; j--) std::swap(r[j - 1], r[j]);  // promote
        return true;
      }
    }
    Cell& last = r[mru - 1];
    last.used = true;
    last.node = n;
    last.addr = 0;
    for (size_t j = mru - 1; j > 0; j--) std::swap(r[j - 1], r[j]);
    *slot = &r[0].addr;
    return false;
  }
};

struct Builder {
  std::vector<uint8_t> out;
  std::vector<Unfinished> stack;
  Registry registry{10000, 2};
  std::string last_key;
  bool any = false;
  uint64_t len = 0;
  uint64_t last_addr = kNoneAddr;

  Builder() {
    out.resize(kHeader, 0);
    out[0] = 1;  // version 1, type 0
    stack.emplace_back();  // root
  }
  void put(uint8_t b) { out.push_back(b); }
  void put_packed(uint64_t v, int n) {
    for (int i = 0; i < n; i++) {
      out.push_back((uint8_t)v);
      v >>= 8;
    }
  }
  static uint64_t delta_addr(uint64_t base, uint64_t trans) { return trans == 0 ? 0 : base - trans; }

  uint64_t encode_one_finish(const Node& s, uint8_t next) {
    const uint8_t enc = kCommon.code[s.trans[0].in];
    if (enc == 0) put(s.trans[0].in);
    put((uint8_t)(kOneTransition | next | enc));
    return out.size() - 1;
  }
  uint64_t encode_one(const Node& s) {
    const uint64_t start = out.size();
    int out_size = 0;
    if (s.trans[0].out != 0) {
      out_size = packed_size(s.trans[0].out);
      put_packed(s.trans[0].out, out_size);
    }
    const uint64_t delta = delta_addr(start, s.trans[0].addr);
    const int trans_size = packed_size(delta);
    put_packed(delta, trans_size);
    put((uint8_t)(trans_size << 4 | out_size));
    return encode_one_finish(s, 0);
  }
  uint64_t encode_many(const Node& s) {
    const uint64_t start = out.size();
    int trans_size = 0, out_size = packed_size(s.final_output);
    bool any_outputs = s.final_output != 0;
    for (const Trans& t : s.trans) {
      trans_size = std::max(trans_size, packed_size(delta_addr(start, t.addr)));
      out_size = std::max(out_size, packed_size(t.out));
      any_outputs = any_outputs || t.out != 0;
    }
    if (!any_outputs) out_size = 0;
    if (any_outputs) {
      if (s.final) put_packed(s.final_output, out_size);
      for (size_t j = s.trans.size(); j-- > 0;) put_packed(s.trans[j].out, out_size);
    }
    for (size_t j = s.trans.size(); j-- > 0;)
      put_packed(delta_addr(start, s.trans[j].addr), trans_size);
    for (size_t j = s.trans.size(); j-- > 0;) put(s.trans[j].in);
    put((uint8_t)(trans_size << 4 | out_size));
    uint8_t num = s.trans.size() <= kMaxNumTrans ? (uint8_t)s.trans.size() : 0;
    if (num == 0) put(s.trans.size() == 256 ? (uint8_t)1 : (uint8_t)s.trans.size());
    if (s.final) num |= kStateFinal;
    put(num);
    return out.size() - 1;
  }
  uint64_t encode_state(const Node& s) {
    if (s.trans.empty() && s.final && s.final_output == 0) return kEmptyAddr;
    if (s.trans.size() != 1 || s.final) return encode_many(s);
    if (s.trans[0].out == 0 && s.trans[0].addr == last_addr)
      return encode_one_finish(s, kTransitionNext);
    return encode_one(s);
  }
  uint64_t compile(const Node& n) {
    if (n.final && n.trans.empty() && n.final_output == 0) return kEmptyAddr;
    uint64_t addr = 0, *slot = nullptr;
    if (registry.entry(n, addr, &slot)) return addr;
    addr = encode_state(n);
    last_addr = addr;
    *slot = addr;
    return addr;
  }
  void compile_from(size_t i_state) {
    uint64_t addr = kNoneAddr;
    while (i_state + 1 < stack.size()) {
      Unfinished u = std::move(stack.back());
      stack.pop_back();
      if (addr != kNoneAddr) u.last_compiled(addr);
      addr = compile(u.node);
    }
    stack.back().last_compiled(addr);
  }
  // II2_OK or II2_ERR_INVALID (keys must ascend strictly)
  int insert(const uint8_t* key, size_t n, uint64_t val) {
    if (any) {
      const size_t m = std::min(n, last_key.size());
      int c = m ? memcmp(key, last_key.data(), m) : 0;
      if (c == 0) c = n < last_key.size() ? -1 : (n > last_key.size() ? 1 : 0);
      if (c <= 0) return II2_ERR_INVALID;
    }
    any = true;
    if (n == 0) {
      len = 1;
      stack[0].node.final = true;
      stack[0].node.final_output = val;
      last_key.clear();
      return II2_OK;
    }
    // common prefix with the previous key; outputs are pushed down so that every prefix
    // carries the minimum of the values below it
    size_t i = 0;
    uint64_t outv = val;
    while (i < n && i < stack.size() && stack[i].has_last && stack[i].last_in == key[i]) {
      const uint64_t common = std::min(stack[i].last_out, outv);
      const uint64_t add = stack[i].last_out - common;
      outv -= common;
      stack[i].last_out = common;
      i++;
      if (add != 0) stack[i].add_output_prefix(add);
    }
    len++;
    compile_from(i);
    last_key.assign(reinterpret_cast<const char*>(key), n);
    // the suffix: one unfinished node per remaining byte, then the final empty one
    Unfinished& top = stack.back();
    top.has_last = true;
    top.last_in = key[i];
    top.last_out = outv;
    for (size_t j = i + 1; j < n; j++) {
      Unfinished u;
      u.has_last = true;
      u.last_in = key[j];
      stack.push_back(std::move(u));
    }
    Unfinished fin;
    fin.node.final = true;
    stack.push_back(std::move(fin));
    return II2_OK;
  }
  void close() {
    compile_from(0);
    const Node root = stack[0].node;
    const uint64_t root_addr = compile(root);
    put_packed(len, 8);
    put_packed(root_addr, 8);
  }
};

// ---------------------------------------------------------------- reader (decoder_v1.go)
struct State {
  bool final = false;
  uint64_t final_out = 0;
  uint32_t num = 0;
  // single-transition form
  bool single = false;
  uint8_t s_in = 0;
  uint64_t s_addr = 0, s_out = 0;
  // multi form: regions inside the file
  size_t trans_top = 0, dest_top = 0, out_top = 0, bottom = 0;
  int trans_size = 0, out_size = 0;
};

struct Fst {
  const uint8_t* d;
  size_t n;
  uint64_t len = 0, root = 0;

  int open() {
    if (n < kHeader + kFooter) return II2_ERR_CORRUPT;
    if (read_u64(d) != 1) return II2_ERR_UNSUPPORTED;  // encoding version
    len = read_u64(d + n - 16);
    root = read_u64(d + n - 8);
    return II2_OK;
  }
  // the region of states ends where the footer begins
  size_t limit() const { return n - kFooter; }

  int state_at(uint64_t addr, State& s) const {
    s = State();
    if (addr == kEmptyAddr) {
      s.final = true;
      return II2_OK;
    }
    if (addr < kHeader || addr >= limit()) return II2_ERR_CORRUPT;
    size_t bottom = (size_t)addr;
    const uint8_t hdr = d[addr];
    auto take = [&](size_t k) -> bool {  // move `bottom` down by k bytes
      if (bottom < kHeader + k) return false;
      bottom -= k;
      return true;
    };
    if (hdr & kOneTransition) {
      s.single = true;
      s.num = 1;
      const bool next = (hdr & kTransitionNext) != 0;
      const uint8_t code = hdr & kMaxCommon;
      if (code == 0) {
        if (!take(1)) return II2_ERR_CORRUPT;
        s.s_in = d[bottom];
      } else {
        s.s_in = kCommonInv[code - 1];
      }
      if (next) {
        if (bottom < kHeader + 1) return II2_ERR_CORRUPT;
        s.s_addr = bottom - 1;
        s.s_out = 0;
      } else {
        if (!take(1)) return II2_ERR_CORRUPT;
        const int ts = d[bottom] >> 4, os = d[bottom] & 15;
        if (ts > 8 || os > 8 || !take((size_t)ts)) return II2_ERR_CORRUPT;
        uint64_t delta = read_packed(d + bottom, ts);
        if (os > 0) {
          if (!take((size_t)os)) return II2_ERR_CORRUPT;
          s.s_out = read_packed(d + bottom, os);
        }
        if (delta != 0) {
          if (delta > bottom) return II2_ERR_CORRUPT;
          delta = bottom - delta;
        }
        s.s_addr = delta;
      }
      s.bottom = bottom;
      return II2_OK;
    }
    s.final = (hdr & kStateFinal) != 0;
    s.num = hdr & kMaxNumTrans;
    if (s.num == 0) {
      if (!take(1)) return II2_ERR_CORRUPT;
      s.num = d[bottom];
      if (s.num == 1) s.num = 256;
    }
    if (!take(1)) return II2_ERR_CORRUPT;
    s.trans_size = d[bottom] >> 4;
    s.out_size = d[bottom] & 15;
    if (s.trans_size > 8 || s.out_size > 8) return II2_ERR_CORRUPT;
    s.trans_top = bottom;
    if (!take(s.num)) return II2_ERR_CORRUPT;
    s.dest_top = bottom;
    if (!take((size_t)s.num * s.trans_size)) return II2_ERR_CORRUPT;
    if (s.out_size > 0) {
      s.out_top = bottom;
      if (!take((size_t)s.num * s.out_size)) return II2_ERR_CORRUPT;
      if (s.final) {
        if (!take((size_t)s.out_size)) return II2_ERR_CORRUPT;
        s.final_out = read_packed(d + bottom, s.out_size);
      }
    }
    s.bottom = bottom;
    return II2_OK;
  }
  // transition i (ascending input order) of a state
  int transition(const State& s, uint32_t i, uint8_t& in, uint64_t& addr, uint64_t& out) const {
    if (s.single) {
      in = s.s_in;
      addr = s.s_addr;
      out = s.s_out;
      return II2_OK;
    }
    in = d[s.trans_top - i - 1];
    const uint64_t delta = read_packed(d + s.dest_top - (size_t)(i + 1) * s.trans_size, s.trans_size);
    if (delta > s.bottom) return II2_ERR_CORRUPT;
    addr = delta ? s.bottom - delta : 0;
    out = s.out_size ? read_packed(d + s.out_top - (size_t)(i + 1) * s.out_size, s.out_size) : 0;
    return II2_OK;
  }
};

struct TermsOwner {
  std::vector<uint8_t> tb;
  std::vector<uint32_t> off;
  std::vector<uint64_t> val;
};

int bytes_compare(const uint8_t* a, size_t na, const uint8_t* b, size_t nb) {
  const size_t m = std::min(na, nb);
  const int c = m ? memcmp(a, b, m) : 0;
  if (c) return c;
  return na < nb ? -1 : (na > nb ? 1 : 0);
}

}  // namespace

extern "C" {

int ii2_fst_build(const uint8_t* term_bytes, const uint32_t* term_off, const uint64_t* values,
                  uint64_t n_terms, uint8_t** fst, uint64_t* nbytes) {
  if (!fst || !nbytes || (n_terms && (!term_off || !values))) return II2_ERR_INVALID;
  *fst = nullptr;
  *nbytes = 0;
  try {
    Builder b;
    for (uint64_t i = 0; i < n_terms; i++) {
      if (term_off[i + 1] < term_off[i]) return II2_ERR_INVALID;
      const int rc = b.insert(term_bytes + term_off[i], term_off[i + 1] - term_off[i], values[i]);
      if (rc != II2_OK) return rc;
    }
    b.close();
    uint8_t* p = static_cast<uint8_t*>(malloc(b.out.size()));
    if (!p) return II2_ERR_NOMEM;
    memcpy(p, b.out.data(), b.out.size());
    *fst = p;
    *nbytes = b.out.size();
  } catch (const std::bad_alloc&) {
    return II2_ERR_NOMEM;
  }
  return II2_OK;
}

void ii2_fst_free(void* p) { free(p); }

int ii2_fst_len(const uint8_t* fst, uint64_t nbytes, uint64_t* n_terms) {
  if (!fst || !n_terms) return II2_ERR_INVALID;
  Fst f{fst, (size_t)nbytes};
  const int rc = f.open();
  if (rc != II2_OK) return rc;
  *n_terms = f.len;
  return II2_OK;
}

int ii2_fst_get(const uint8_t* fst, uint64_t nbytes, const uint8_t* key, size_t keylen,
                uint64_t* value, int* found) {
  if (!fst || !value || !found || (keylen && !key)) return II2_ERR_INVALID;
  *found = 0;
  *value = 0;
  Fst f{fst, (size_t)nbytes};
  int rc = f.open();
  if (rc != II2_OK) return rc;
  State s;
  if ((rc = f.state_at(f.root, s)) != II2_OK) return rc;
  uint64_t total = 0;
  for (size_t i = 0; i < keylen; i++) {
    bool hit = false;
    for (uint32_t t = 0; t < s.num; t++) {
      uint8_t in;
      uint64_t addr, out;
      if ((rc = f.transition(s, t, in, addr, out)) != II2_OK) return rc;
      if (in == key[i]) {
        total += out;
        if ((rc = f.state_at(addr, s)) != II2_OK) return rc;
        hit = true;
        break;
      }
    }
    if (!hit) return II2_OK;
  }
  if (s.final) {
    *found = 1;
    *value = total + s.final_out;
  }
  return II2_OK;
}

int ii2_fst_read(const uint8_t* fst, uint64_t nbytes, const uint8_t* min, size_t minlen,
                 const uint8_t* max, size_t maxlen, ii2_fst_terms* out) {
  if (!fst || !out) return II2_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  Fst f{fst, (size_t)nbytes};
  int rc = f.open();
  if (rc != II2_OK) return rc;
  try {
    TermsOwner* own = new TermsOwner();
    struct Guard {
      TermsOwner* p;
      ~Guard() { delete p; }
    } guard{own};
    own->off.push_back(0);
    struct Frame {
      State st;
      uint32_t next;
      uint64_t sum;  // outputs accumulated on the way to this state
      bool on_min;   // the key so far equals min[0 .. depth)
    };
    std::vector<Frame> stack;
    std::vector<uint8_t> key;
    const bool has_min = min != nullptr, has_max = max != nullptr;
    bool done = false;
    // a state is entered: emit its key if it is final and inside the bounds
    auto enter = [&](uint64_t addr, uint64_t sum, bool on_min) -> int {
      Frame fr;
      const int r = f.state_at(addr, fr.st);
      if (r != II2_OK) return r;
      fr.next = 0;
      fr.sum = sum;
      fr.on_min = on_min;
      if (fr.st.final) {
        // on the min path a key shorter than min is a proper prefix of it: smaller
        const bool ge_min = !on_min || key.size() >= minlen;
        if (ge_min) {
          if (has_max && bytes_compare(key.data(), key.size(), max, maxlen) > 0) {
            done = true;  // ascending order: nothing after this key is in range
            return II2_OK;
          }
          if (own->tb.size() + key.size() > 0xFFFFFFFFull) return II2_ERR_UNSUPPORTED;
          own->tb.insert(own->tb.end(), key.begin(), key.end());
          own->off.push_back((uint32_t)own->tb.size());
          own->val.push_back(sum + fr.st.final_out);
        }
      }
      if (key.size() > 65535 + 1) return II2_ERR_CORRUPT;  // a cycle: no term is that long
      stack.push_back(fr);
      return II2_OK;
    };
    if ((rc = enter(f.root, 0, has_min)) != II2_OK) return rc;
    while (!stack.empty() && !done) {
      Frame& fr = stack.back();
      if (fr.next >= fr.st.num) {
        stack.pop_back();
        if (!key.empty()) key.pop_back();
        continue;
      }
      const uint32_t t = fr.next++;
      uint8_t in;
      uint64_t addr, o;
      if ((rc = f.transition(fr.st, t, in, addr, o)) != II2_OK) return rc;
      bool child_on_min = false;
      const size_t depth = stack.size() - 1;  // bytes of the key before this transition
      if (fr.on_min && depth < minlen) {
        if (in < min[depth]) continue;  // everything below sorts before min
        child_on_min = in == min[depth];
      }
      const uint64_t sum = fr.sum + o;
      key.push_back(in);
      if ((rc = enter(addr, sum, child_on_min)) != II2_OK) return rc;
      if (done) break;
    }
    own->tb.resize(own->tb.size() + 32, 0);  // readable padding for over-reading key loads
    out->n_terms = own->val.size();
    out->term_bytes = own->tb.data();
    out->term_off = own->off.data();
    out->values = own->val.data();
    out->fst_len = f.len;
    out->_owner = own;
    guard.p = nullptr;
  } catch (const std::bad_alloc&) {
    return II2_ERR_NOMEM;
  }
  return II2_OK;
}

void ii2_fst_terms_free(ii2_fst_terms* t) {
  if (!t) return;
  delete static_cast<TermsOwner*>(t->_owner);
  memset(t, 0, sizeof(*t));
}

}  // extern "C"
