module go_vectors

go 1.22.5

// the reference under test and the versions its own go.mod pins (go.mod:5-13)
require (
	github.com/RoaringBitmap/roaring v1.9.4
	github.com/blevesearch/vellum v1.0.10
	github.com/lezhnev74/inverted_index_2 v0.0.0
	github.com/ronanh/intcomp v1.1.0
)

// point this at a checkout of the reference (the commit being replaced)
replace github.com/lezhnev74/inverted_index_2 => ../../../reference
