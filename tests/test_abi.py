"""The C-ABI library loads and exports every symbol include/gsrast_b200.h declares; the
host-only entry points (size probes, chunk carving, msb) behave like the reference's
required<T> / fromChunk (AuxBuffer.cuh:8-14, AuxBuffer.cu:13-89).  No compute calls: runs on CPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from gsrast_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib.lib()


def test_header_symbols_exported(L):
    from gsrast_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "gsrast_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gsr_[a-z0-9_]+)\s*\(", hdr)) - {"gsr_alloc_fn"}
    assert len(declared) >= 20
    for sym in sorted(declared):
        assert hasattr(L, sym), "libgsrast_b200.so does not export %s" % sym
    assert declared == set(_lib.EXPORTS)
    assert L.gsr_version() == 200


def test_cpp_shim_headers_present():
    for h in ("rasterizer.h", "gscuda_dropin.h", "gsrast_b200.h"):
        assert os.path.exists(os.path.join(ROOT, "include", h))


def test_required_sizes_and_alignment(L):
    from gsrast_b200 import _lib

    prev = 0
    for P in (0, 1, 255, 256, 257, 100_000, 3_300_000, 6_000_000):
        n = L.gsr_geometry_state_required(P)
        assert n >= prev
        prev = n
        st = _lib.GeometryState()
        base = 1 << 20
        used = L.gsr_geometry_state_map(base + 4, P, C.byref(st))  # deliberately misaligned chunk
        assert used + 4 <= n
        for name, ctype in st._fields_:
            if ctype is C.c_size_t:
                continue
            v = getattr(st, name)
            for addr in (list(v) if hasattr(v, "__len__") else [v]):
                addr = addr or 0
                assert addr % 128 == 0 and addr >= base, name
    # geometry scratch per Gaussian: the reference's 79 B (+ our block sums) plus the depth half of the
    # radix sort (depth keys 4 B + tile rect 8 B + two key/id ping-pong pairs 16 B + look-back state 2 B)
    assert L.gsr_geometry_state_required(3_300_000) / 3.3e6 < 119
    prev = 0
    for R in (0, 1, 4096, 4097, 15_000_000, 60_000_000):
        n = L.gsr_binning_state_required(R)
        assert n >= prev and n >= 24 * R
        prev = n
        st = _lib.BinningState()
        L.gsr_binning_state_map(1 << 20, R, C.byref(st))
        assert st.point_list_keys - st.point_list_keys_unsorted >= 8 * R
        assert st.sorting_size >= L.gsr_sort_pairs_temp_bytes(R) + 4 * R  # second id ping-pong array
    # image state: ranges per TILE (the reference clears W*H entries, GSCuda.cu:800)
    assert L.gsr_image_state_required(1920, 1080) < 8 * 1920 * 1080 + 3 * 8160 * 8 + 1024
    st = _lib.ImageState()
    L.gsr_image_state_map(1 << 20, 1920, 1080, C.byref(st))
    assert st.n_contrib - st.ranges >= 8160 * 8


def test_error_strings(L):
    from gsrast_b200 import _lib

    assert b"invalid" in L.gsr_error_string(_lib.ERR_INVALID_ARG)
    assert L.gsr_error_string(0) == b"success"
    assert L.gsr_forward_ex(None) == _lib.ERR_INVALID_ARG  # argument validation happens before any CUDA call


def test_msb(L):
    for n, want in [(3072, 12), (3600, 12), (8160, 13), (32400, 15)]:
        assert L.gsr_get_higher_msb(n) == want


def test_product_never_imports_oracle():
    """The product path must not import, link or call the oracle (or any CPU fallback)."""
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b|libgsr_oracle|gsr_oracle_\w+\s*\(|from\s+\.\.?oracle)", re.M)
    for top in ("gsrast_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(src), "%s uses the oracle" % os.path.join(dirpath, f)
    # and the shared library does not link it
    import subprocess

    from gsrast_b200 import _lib
    out = subprocess.check_output(["ldd", _lib.LIB_PATH]).decode()
    assert "gsr_oracle" not in out
