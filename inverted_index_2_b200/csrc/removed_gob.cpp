// removed_gob.cpp — the `removed.list` file of a shard: encoding/gob stream of a
// map[int64][]uint32 (RemovedLists.Serialize / UnserializeRemovedList, removed_list.go:26-33,
// 73-80; written next to the segment files at shard.go:340-358).  SURVEY 8f row 4: with this and
// fst_v1.cpp a native process and the Go index can share a shard directory.
//
// Host-side C++, no device work.  PARITY UNPINNED: Go's standard library is not in this image;
// the stream layout below restates the published gob wire format ("encoding/gob" package
// documentation): every message is a uint byte count followed by a signed type id and a body;
// a negative id introduces the wireType description of type -id; unsigned integers below 128
// are one byte, otherwise a byte holding the negated byte count followed by the big-endian
// value; signed integers are (i << 1) or (^i << 1 | 1); a struct is a sequence of (field-number
// delta, value) pairs closed by a zero; a top-level non-struct value is preceded by a zero
// byte; a map is a count followed by key/value pairs; a slice is a count followed by its elements.
// The decoder walks ANY well-formed stream whose value is a map from a signed integer to a slice
// of unsigned integers (type ids and optional type names are taken from the stream); the
// encoder writes the two type descriptions in the order Go sends them (the map, then its
// element slice) with the ids a fresh Go process would assign (slice 65, map 66) and without
// type names, which gob's decoder does not need.  Map iteration order in Go is random, so byte
// equality with a Go-written file is not defined even in principle; lists are written in
// ascending timestamp order here.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <new>
#include <vector>

#include "../../include/ii2.h"

namespace {

struct Writer {
  std::vector<uint8_t> b;
  void u(uint64_t v) {
    if (v < 128) {
      b.push_back((uint8_t)v);
      return;
    }
    int n = 8;
    while (n > 1 && (v >> (8 * (n - 1))) == 0) n--;
    b.push_back((uint8_t)(256 - n));
    for (int i = n - 1; i >= 0; i--) b.push_back((uint8_t)(v >> (8 * i)));
  }
  void i(int64_t v) {
    const uint64_t x = v < 0 ? ((~(uint64_t)v) << 1) | 1 : ((uint64_t)v << 1);
    u(x);
  }
};

void message(std::vector<uint8_t>& out, const Writer& body) {
  Writer len;
  len.u(body.b.size());
  out.insert(out.end(), len.b.begin(), len.b.end());
  out.insert(out.end(), body.b.begin(), body.b.end());
}

constexpr int64_t kTInt = 2, kTUint = 3;        // gob's built-in type ids
constexpr int64_t kSliceId = 65, kMapId = 66;   // first ids a Go process hands to user types

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  uint64_t u() {
    if (p >= end) return fail();
    const uint8_t c = *p++;
    if (c < 128) return c;
    const int n = 256 - c;
    if (n > 8 || end - p < n) return fail();
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v = (v << 8) | *p++;
    return v;
  }
  int64_t i() {
    const uint64_t x = u();
    return (x & 1) ? (int64_t)~(x >> 1) : (int64_t)(x >> 1);
  }
  uint64_t fail() {
    ok = false;
    p = end;
    return 0;
  }
  void skip(uint64_t n) {
    if ((uint64_t)(end - p) < n) {
      fail();
      return;
    }
    p += n;
  }
};

struct TypeDesc {
  int kind = 0;  // 1 = slice, 2 = map
  int64_t key = 0, elem = 0;
};

// CommonType { Name string; Id typeId }: both optional on the wire
void read_common(Reader& r) {
  int field = -1;
  for (;;) {
    const uint64_t d = r.u();
    if (!r.ok || d == 0) return;
    field += (int)d;
    if (field == 0)
      r.skip(r.u());  // Name
    else if (field == 1)
      r.i();  // Id
    else {
      r.fail();
      return;
    }
  }
}

// sliceType { CommonType; Elem } / mapType { CommonType; Key; Elem }
void read_composite(Reader& r, TypeDesc& t, bool is_map) {
  int field = -1;
  for (;;) {
    const uint64_t d = r.u();
    if (!r.ok || d == 0) return;
    field += (int)d;
    if (field == 0)
      read_common(r);
    else if (field == 1)
      (is_map ? t.key : t.elem) = r.i();
    else if (field == 2 && is_map)
      t.elem = r.i();
    else {
      r.fail();
      return;
    }
  }
}

struct Lists {
  std::vector<int64_t> ts;
  std::vector<uint64_t> off;
  std::vector<uint32_t> val;
};

}  // namespace

extern "C" {

int ii2_removed_list_encode(const int64_t* timestamps, const uint64_t* off, const uint32_t* values,
                            uint64_t n_lists, uint8_t** bytes, uint64_t* nbytes) {
  if (!bytes || !nbytes || (n_lists && (!timestamps || !off))) return II2_ERR_INVALID;
  *bytes = nullptr;
  *nbytes = 0;
  try {
    std::map<int64_t, uint64_t> order;  // ascending timestamps, the last duplicate wins
    for (uint64_t k = 0; k < n_lists; k++) {
      if (off[k + 1] < off[k]) return II2_ERR_INVALID;
      order[timestamps[k]] = k;
    }
    std::vector<uint8_t> out;
    {  // type 66 = map[int64][]uint32: wireType{MapT: {CommonType{Id: 66}, Key: int, Elem: 65}}
      Writer w;
      w.i(-kMapId);
      w.u(4);  // wireType field 3 (MapT), from -1
      w.u(1);  // mapType field 0: CommonType
      w.u(2);  // CommonType field 1 (Id); Name omitted
      w.i(kMapId);
      w.u(0);
      w.u(1);  // Key
      w.i(kTInt);
      w.u(1);  // Elem
      w.i(kSliceId);
      w.u(0);  // end mapType
      w.u(0);  // end wireType
      message(out, w);
    }
    {  // type 65 = []uint32: wireType{SliceT: {CommonType{Id: 65}, Elem: uint}}
      Writer w;
      w.i(-kSliceId);
      w.u(2);  // wireType field 1 (SliceT)
      w.u(1);
      w.u(2);
      w.i(kSliceId);
      w.u(0);
      w.u(1);  // Elem
      w.i(kTUint);
      w.u(0);
      w.u(0);
      message(out, w);
    }
    {
      Writer w;
      w.i(kMapId);
      w.u(0);  // singleton: the value is not a struct
      w.u(order.size());
      for (const auto& kv : order) {
        const uint64_t k = kv.second;
        w.i(kv.first);
        w.u(off[k + 1] - off[k]);
        for (uint64_t e = off[k]; e < off[k + 1]; e++) w.u(values[e]);
      }
      message(out, w);
    }
    uint8_t* p = static_cast<uint8_t*>(malloc(out.size()));
    if (!p) return II2_ERR_NOMEM;
    memcpy(p, out.data(), out.size());
    *bytes = p;
    *nbytes = out.size();
  } catch (const std::bad_alloc&) {
    return II2_ERR_NOMEM;
  }
  return II2_OK;
}

int ii2_removed_list_decode(const uint8_t* bytes, uint64_t nbytes, ii2_removed_lists* out) {
  if (!out || (nbytes && !bytes)) return II2_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  try {
    Lists* L = new Lists();
    struct Guard {
      Lists* p;
      ~Guard() { delete p; }
    } guard{L};
    L->off.push_back(0);
    std::map<int64_t, TypeDesc> types;
    Reader top{bytes, bytes + nbytes};
    bool have_value = false;
    while (top.p < top.end && !have_value) {
      const uint64_t len = top.u();
      if (!top.ok || (uint64_t)(top.end - top.p) < len) return II2_ERR_CORRUPT;
      Reader r{top.p, top.p + len};
      top.p += len;
      const int64_t id = r.i();
      if (!r.ok) return II2_ERR_CORRUPT;
      if (id < 0) {  // wireType of type -id: exactly one of its pointer fields is set
        TypeDesc t;
        int field = -1;
        for (;;) {
          const uint64_t d = r.u();
          if (!r.ok) return II2_ERR_CORRUPT;
          if (d == 0) break;
          field += (int)d;
          if (field == 1) {
            t.kind = 1;
            read_composite(r, t, false);
          } else if (field == 3) {
            t.kind = 2;
            read_composite(r, t, true);
          } else {
            return II2_ERR_UNSUPPORTED;  // array / struct / GobEncoder: not a removed list
          }
          if (!r.ok) return II2_ERR_CORRUPT;
        }
        types[-id] = t;
        continue;
      }
      const auto mt = types.find(id);
      if (mt == types.end() || mt->second.kind != 2 || mt->second.key != kTInt)
        return II2_ERR_UNSUPPORTED;
      const auto st = types.find(mt->second.elem);
      if (st == types.end() || st->second.kind != 1 || st->second.elem != kTUint)
        return II2_ERR_UNSUPPORTED;
      if (r.u() != 0 || !r.ok) return II2_ERR_CORRUPT;  // singleton marker
      const uint64_t n = r.u();
      for (uint64_t k = 0; k < n && r.ok; k++) {
        L->ts.push_back(r.i());
        const uint64_t m = r.u();
        if (!r.ok || m > (uint64_t)(r.end - r.p)) return II2_ERR_CORRUPT;  // >= 1 byte per value
        for (uint64_t e = 0; e < m && r.ok; e++) {
          const uint64_t v = r.u();
          if (v > 0xFFFFFFFFull) return II2_ERR_CORRUPT;
          L->val.push_back((uint32_t)v);
        }
        L->off.push_back(L->val.size());
      }
      if (!r.ok || r.p != r.end) return II2_ERR_CORRUPT;
      have_value = true;
    }
    if (!have_value && nbytes) return II2_ERR_CORRUPT;
    out->n_lists = L->ts.size();
    out->timestamps = L->ts.data();
    out->off = L->off.data();
    out->values = L->val.data();
    out->_owner = L;
    guard.p = nullptr;
  } catch (const std::bad_alloc&) {
    return II2_ERR_NOMEM;
  }
  return II2_OK;
}

void ii2_removed_lists_free(ii2_removed_lists* l) {
  if (!l) return;
  delete static_cast<Lists*>(l->_owner);
  memset(l, 0, sizeof(*l));
}

}  // extern "C"
