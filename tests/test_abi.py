"""CPU-only checks of the drop-in boundary: libii2.so loads, exports every symbol include/ii2.h
declares, and refuses to compute without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from inverted_index_2_b200 import _abi as A
from inverted_index_2_b200.engine import load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ii2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ii2_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    return load_library()


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libii2.so does not export {n}"
    assert sorted(A.PROTOTYPES) == names, "ctypes prototypes and include/ii2.h disagree"


def test_abi_version_and_strerror(lib):
    assert lib.ii2_abi_version() == 2
    assert b"bitmask is out of bound" in lib.ii2_strerror(A.II2_ERR_BITMASK_OOB)
    assert b"no CPU fallback" in lib.ii2_strerror(A.II2_ERR_NO_DEVICE)


def test_shard_key_host_helper(lib):
    for term, key in [(b"", 0), (b"a", 0), (b"ab", (97 * 256 + 98) >> 6), (b"\xff\xff", 1023)]:
        buf = (C.c_uint8 * max(1, len(term))).from_buffer_copy(term.ljust(1, b"\0"))
        assert lib.ii2_shard_key(C.cast(buf, A.u8p), len(term)) == key


def test_no_cpu_fallback(lib):
    """Without a device the compute entry points fail loudly with II2_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    assert lib.ii2_init(None, 0) == A.II2_ERR_NO_DEVICE
    out = A.MergeOut()
    assert lib.ii2_merge(None, 0, None, 0, 0, C.byref(out)) == A.II2_ERR_NO_DEVICE
    rout = A.ReadOut()
    assert lib.ii2_read_range(None, 0, None, 0, None, 0, None, 0, C.byref(rout)) == A.II2_ERR_NO_DEVICE
    w, o = A.u32p(), A.u64p()
    off = np.zeros(1, dtype=np.uint64)
    assert lib.ii2_intcomp_encode_u32(None, A.np_ptr(off, A.u64p), 0, C.byref(w), C.byref(o)) \
        == A.II2_ERR_NO_DEVICE
    h = C.c_void_p()
    assert lib.ii2_bitmask_new(None, 0, C.byref(h)) == A.II2_ERR_NO_DEVICE
    from inverted_index_2_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)
