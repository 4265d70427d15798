// go_vectors — writes tests/golden/go_vectors.json: bytes produced by the REAL Go code of the
// reference (github.com/lezhnev74/inverted_index_2 at the commit under test) and by the pinned
// third-party modules of its go.mod (ronanh/intcomp v1.1.0, blevesearch/vellum v1.0.10,
// RoaringBitmap/roaring v1.9.4, encoding/gob), for the inputs the B200 library's tests use.
//
// This image has no Go toolchain, so this program has never been compiled here.  Run it on
// any machine with Go >= 1.22 and network access to the module proxy:
//
//	cd scripts/go_vectors
//	go mod tidy          # resolves the versions pinned in go.mod (same as the reference's)
//	go run . > ../../tests/golden/go_vectors.json
//
// tests/test_go_vectors.py then compares every byte with the oracle (CPU) and the CUDA library
// (GPU) and fails on the first difference; while the file is absent those tests SKIP and the
// formats stay "parity unpinned".
package main

import (
	"bytes"
	"encoding/gob"
	"encoding/hex"
	"encoding/json"
	"fmt"
	"math/rand"
	"os"
	"path/filepath"

	"github.com/RoaringBitmap/roaring"
	"github.com/blevesearch/vellum"
	"github.com/lezhnev74/inverted_index_2/file"
	"github.com/ronanh/intcomp"
)

type intcompCase struct {
	Name   string   `json:"name"`
	Values []uint32 `json:"values"`
	Words  []uint32 `json:"words"` // intcomp.CompressUint32(values, nil)
}

type roaringCase struct {
	Name   string   `json:"name"`
	DictN  int      `json:"dict_n"` // dictionary = 0 .. dict_n-1 (0: grow on miss)
	Puts   [][]uint32 `json:"puts"`
	Hex    []string `json:"hex"` // Bitmask.Put bytes, one per put
	Values []uint32 `json:"all_values"`
}

type item struct {
	Term   string   `json:"term"`
	Values []uint32 `json:"values"`
}

type segmentCase struct {
	Name   string `json:"name"`
	Direct bool   `json:"direct"`
	Items  []item `json:"items"`
	FstHex string `json:"fst_hex"` // bytes of <key>_fst written by file.Writer
	ValHex string `json:"val_hex"` // bytes of <key>_val ("" in direct mode)
}

type fstCase struct {
	Name   string   `json:"name"`
	Keys   []string `json:"keys_hex"`
	Vals   []uint64 `json:"values"`
	FstHex string   `json:"fst_hex"` // vellum.New(w, nil) + Insert* + Close
}

type gobCase struct {
	Name  string             `json:"name"`
	Lists map[string][]uint32 `json:"lists"` // key = decimal timestamp
	Hex   string             `json:"hex"`   // gob.NewEncoder(buf).Encode(map[int64][]uint32)
}

type output struct {
	Modules  map[string]string `json:"modules"`
	Intcomp  []intcompCase     `json:"intcomp"`
	Roaring  []roaringCase     `json:"roaring"`
	Segments []segmentCase     `json:"segments"`
	Fst      []fstCase         `json:"fst"`
	Gob      []gobCase         `json:"gob"`
}

func must(err error) {
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
}

func seq(first, n, step uint32) []uint32 {
	out := make([]uint32, n)
	for i := range out {
		out[i] = first + uint32(i)*step
	}
	return out
}

func randomSorted(r *rand.Rand, n int, maxGap uint32) []uint32 {
	out := make([]uint32, n)
	var cur uint32
	for i := range out {
		cur += 1 + uint32(r.Intn(int(maxGap)))
		out[i] = cur
	}
	return out
}

func intcompCases() []intcompCase {
	r := rand.New(rand.NewSource(0x1C0))
	mk := func(name string, v []uint32) intcompCase {
		return intcompCase{Name: name, Values: v, Words: intcomp.CompressUint32(v, nil)}
	}
	unsorted := []uint32{10, 500, 300} // file/writer_test.go:14
	wide := make([]uint32, 200)
	for i := range wide {
		wide[i] = r.Uint32()
	}
	return []intcompCase{
		mk("empty", []uint32{}),
		mk("one_value", []uint32{7}),
		mk("writer_test_unsorted", unsorted),
		mk("writer_test_pair", []uint32{66, 5513}),
		mk("zero_first", []uint32{0, 1, 2}),
		mk("max_value", []uint32{0xFFFFFFFF}),
		mk("five_byte_varbyte", []uint32{0x80000000, 0, 0xFFFFFFFF}),
		mk("n127", seq(5, 127, 3)),
		mk("n128", seq(5, 128, 3)),
		mk("n129", seq(5, 129, 3)),
		mk("n255", randomSorted(r, 255, 1000)),
		mk("n256", randomSorted(r, 256, 1000)),
		mk("n257", randomSorted(r, 257, 1<<20)),
		mk("constant_128", seq(9, 128, 0)),
		mk("negative_delta_block", append(seq(1000, 64, 5), seq(10, 64+10, 7)...)),
		mk("random_u32_200", wide),
		mk("n1000_gap1", seq(0, 1000, 1)),
		mk("n1000_gap4096", randomSorted(r, 1000, 8192)),
	}
}

func roaringCases() []roaringCase {
	run := func(name string, dictN int, puts [][]uint32) roaringCase {
		var init []uint32
		if dictN > 0 {
			init = seq(0, uint32(dictN), 1)
		}
		bm := file.NewBitmask[uint32](init)
		c := roaringCase{Name: name, DictN: dictN, Puts: puts}
		for _, p := range puts {
			b, err := bm.Put(p)
			must(err)
			c.Hex = append(c.Hex, hex.EncodeToString(b))
		}
		if dictN == 0 {
			c.Values = bm.AllValues()
		}
		return c
	}
	return []roaringCase{
		run("bitmask_test_put", 0, [][]uint32{{1, 10, 80}, {9, 10, 11}}), // file/bitmask_test.go:34-52
		run("array_one_container", 16, [][]uint32{{3, 1, 2}}),
		run("array_two_containers", 70000, [][]uint32{{65536, 1, 65537, 69999}}),
		run("array_4096_boundary", 5000, [][]uint32{seq(0, 4096, 1)}),
		run("bitmap_4097", 5000, [][]uint32{seq(0, 4097, 1)}),
		run("bitmap_65535", 65536, [][]uint32{seq(0, 65535, 1)}),
		run("run_full_container", 65536, [][]uint32{seq(0, 65536, 1)}),
		run("run_then_array_3_containers", 200000, [][]uint32{seq(0, 131072+10, 1)}),
		run("run_with_offset_header_4_containers", 300000, [][]uint32{seq(0, 3*65536+5, 1)}),
		run("grow_on_miss_70000", 0, [][]uint32{seq(1000000, 70000, 3)}),
	}
}

func readBoth(dir string) (fst, val []byte) {
	m, _ := filepath.Glob(filepath.Join(dir, "*_fst"))
	if len(m) != 1 {
		must(fmt.Errorf("expected one _fst in %s, found %d", dir, len(m)))
	}
	fst, err := os.ReadFile(m[0])
	must(err)
	v, _ := filepath.Glob(filepath.Join(dir, "*_val"))
	if len(v) == 1 {
		val, err = os.ReadFile(v[0])
		must(err)
	}
	return
}

func segmentCases() []segmentCase {
	write := func(name string, direct bool, items []item) segmentCase {
		dir, err := os.MkdirTemp("", "go_vectors")
		must(err)
		defer os.RemoveAll(dir)
		var w *file.Writer
		if direct {
			w, err = file.NewDirectWriter(dir, nil)
		} else {
			w, err = file.NewWriter(dir, nil)
		}
		must(err)
		for _, it := range items {
			must(w.Append(file.TermValues{Term: []byte(it.Term), Values: it.Values}))
		}
		must(w.Close())
		c := segmentCase{Name: name, Direct: direct, Items: items}
		f, v := readBoth(dir)
		c.FstHex, c.ValHex = hex.EncodeToString(f), hex.EncodeToString(v)
		return c
	}
	r := rand.New(rand.NewSource(0x5E6))
	letters := "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
	var many []item
	seen := map[string]bool{}
	for len(many) < 300 {
		n := 10 + r.Intn(10) // inverted_index_test.go:98-100 randomString(10,20)
		b := make([]byte, n)
		for i := range b {
			b[i] = letters[r.Intn(len(letters))]
		}
		if !seen[string(b)] {
			seen[string(b)] = true
			many = append(many, item{Term: string(b), Values: randomSorted(r, 1+r.Intn(200), 5000)})
		}
	}
	// Writer.Append needs ascending terms (vellum): sort
	for i := range many {
		for j := i + 1; j < len(many); j++ {
			if many[j].Term < many[i].Term {
				many[i], many[j] = many[j], many[i]
			}
		}
	}
	return []segmentCase{
		// file/writer_test.go:13-17
		write("TestWriter", false, []item{{"term1", []uint32{10, 500, 300}}, {"term2", []uint32{}}, {"term3", []uint32{66, 5513}}}),
		// file/writer_test.go:52-55
		write("TestWriterDirect", true, []item{{"term1", []uint32{10}}, {"term2", []uint32{11}}}),
		write("shared_prefixes_and_suffixes", false, []item{
			{"a", []uint32{1}}, {"ab", []uint32{1, 2}}, {"abc", []uint32{3}}, {"b", []uint32{4}},
			{"bbc", []uint32{5, 6, 7}}, {"cbc", []uint32{8}}, {"term~", []uint32{9}}}),
		write("random_300_terms", false, many),
	}
}

func fstCases() []fstCase {
	build := func(name string, keys [][]byte, vals []uint64) fstCase {
		var buf bytes.Buffer
		b, err := vellum.New(&buf, nil)
		must(err)
		c := fstCase{Name: name, Vals: vals}
		for i, k := range keys {
			must(b.Insert(k, vals[i]))
			c.Keys = append(c.Keys, hex.EncodeToString(k))
		}
		must(b.Close())
		c.FstHex = hex.EncodeToString(buf.Bytes())
		return c
	}
	all256 := make([][]byte, 256)
	v256 := make([]uint64, 256)
	for i := range all256 {
		all256[i] = []byte{byte(i)}
		v256[i] = uint64(i) * 3
	}
	return []fstCase{
		build("empty", nil, nil),
		build("t_0", [][]byte{[]byte("t")}, []uint64{0}),
		build("tilde_5", [][]byte{[]byte("~")}, []uint64{5}),
		build("empty_key_0", [][]byte{{}}, []uint64{0}),
		build("ab1_ac2", [][]byte{[]byte("ab"), []byte("ac")}, []uint64{1, 2}),
		build("root_256_transitions", all256, v256),
		build("large_outputs", [][]byte{[]byte("x"), []byte("xy"), []byte("z")}, []uint64{1 << 40, (1 << 63) + 5, 0xFFFFFFFFFFFFFFFF}),
	}
}

func gobCases() []gobCase {
	enc := func(name string, lists map[int64][]uint32) gobCase {
		var buf bytes.Buffer
		must(gob.NewEncoder(&buf).Encode(lists))
		c := gobCase{Name: name, Lists: map[string][]uint32{}, Hex: hex.EncodeToString(buf.Bytes())}
		for k, v := range lists {
			c.Lists[fmt.Sprint(k)] = v
		}
		return c
	}
	return []gobCase{
		enc("empty_map", map[int64][]uint32{}),
		enc("one_batch", map[int64][]uint32{1724925600000000001: {1, 5, 10}}),
		enc("removed_list_test", map[int64][]uint32{1: {1, 5, 10}, 2: {2, 20, 30}}), // removed_list_test.go:9-18 (map order is random)
		enc("negative_and_large", map[int64][]uint32{-7: {0xFFFFFFFF}, 1 << 62: {}}),
	}
}

func main() {
	out := output{
		Modules: map[string]string{
			"github.com/ronanh/intcomp":        "v1.1.0",
			"github.com/blevesearch/vellum":    "v1.0.10",
			"github.com/RoaringBitmap/roaring": "v1.9.4",
		},
		Intcomp:  intcompCases(),
		Roaring:  roaringCases(),
		Segments: segmentCases(),
		Fst:      fstCases(),
		Gob:      gobCases(),
	}
	_ = roaring.New // the module is pinned even though only file.Bitmask calls it
	e := json.NewEncoder(os.Stdout)
	e.SetIndent("", " ")
	must(e.Encode(out))
}
